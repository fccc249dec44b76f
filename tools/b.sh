#!/bin/bash
# build the library from anywhere
cd /root/repo && python -m vista_slam_b200.build "$@"
