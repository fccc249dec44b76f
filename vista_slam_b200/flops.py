"""Algorithmic FLOPs of the STA forward pass per image pair (verified there against
torch.utils.flop_counter.FlopCounterMode on the reference forward: 435.8 GF @224x224, 1857.5 GF @512x384)."""


def attention_flops_per_pair(H, W):
    h, w = H // 16, W // 16
    N, M = h * w, h * w + 1
    return 2 * 24 * 4 * N * N * 1024 + 24 * 8 * M * M * 768  # 2 encodes x 24 blocks + 12 x (self + cross) x 2 views


def flops_per_pair(H, W):
    h, w = H // 16, W // 16
    N = h * w
    M = N + 1
    P4 = ((h + 1) // 2) * ((w + 1) // 2)
    enc = 2 * N * 768 * 1024 + 24 * (24 * N * 1024 ** 2 + 4 * N * N * 1024)
    dec = 2 * (2 * N * 1024 * 768) + 24 * (32 * M * 768 ** 2 + 8 * M * M * 768)
    act = (2 * N * 1024 * 96 + 32 * N * 96 ** 2 + 2 * N * 768 * 192 + 8 * N * 192 ** 2 + 2 * N * 768 * 384 +
           2 * N * 768 ** 2 + 18 * P4 * 768 ** 2)
    rn = 18 * 256 * (16 * N * 96 + 4 * N * 192 + N * 384 + P4 * 768)
    refine = (2 * P4 + 4 * N + 16 * N + 64 * N) * 2 * 9 * 256 ** 2 + (4 * P4 + 4 * N + 16 * N + 64 * N) * 2 * 256 ** 2
    head = 2 * 64 * N * 9 * 256 * 128 + 2 * 256 * N * 9 * 128 ** 2 + 2 * 256 * N * 128 * 4
    pose = 2 * (768 * 512 + 2 * 512 ** 2 + 512 * 13)
    return 2 * enc + dec + 2 * (act + rn + refine + head) + 2 * pose
