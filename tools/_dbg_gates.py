import sys, torch
sys.path.insert(0, "/root/repo")
from sgg_b200.engine import Engine
B, T, V = 256, 3, 2000
eng = Engine(B, T, V)
eng.g.init_reference(1); eng.d.init_reference(2)
g = torch.Generator().manual_seed(0)
eng.set_batch(torch.randn(B, 196, 512, generator=g).bfloat16().cuda(), torch.randn(B, 196, 512, generator=g).bfloat16().cuda(), torch.randint(0, V, (B, T), generator=g).cuda())
eng.noise.normal_(); eng.gp_alpha.uniform_()
for i in range(2):
    eng.disc_step(); torch.cuda.synchronize(); print("---- disc step done", flush=True)
