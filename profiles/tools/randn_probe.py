import torch, time
dev = torch.device("cuda:0")
m, k = 61859140, 64
gen = torch.Generator(device=dev); gen.manual_seed(42)
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.time()
    R = torch.randn((m, k), dtype=torch.float64, device=dev, generator=gen) / (k ** 0.5)
    torch.cuda.synchronize(); print("randn fp64 [m,64] + scale:", round((time.time() - t0) * 1e3, 1), "ms")
    del R
