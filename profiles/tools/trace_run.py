import sys, torch
sys.path.insert(0, '.')
from drone_b200.vec import RaceVec
n = 1 << 20
vec = RaceVec(n, seed=0)
g = torch.Generator(device='cpu').manual_seed(1234)
tape = (torch.rand((16, n, 4), generator=g) * 2 - 1).cuda()
vec.reset(0)
vec.step_tape(tape, 0, 300)
torch.cuda.synchronize()
vec.close()
