"""One small CTA-pair GEMM launch (shipped static mode) for compute-sanitizer --tool racecheck: which hazards does the tool
report for the pair kernel as such?"""
import sys
import torch
sys.path.insert(0, ".")
from pixparse_b200 import ops
A = torch.randn((512, 128), device="cuda").bfloat16()
B = torch.randn((256, 128), device="cuda").bfloat16()
out = ops.gemm(A, B, epi=ops.EPI_STORE_BF16)
torch.cuda.synchronize()
print("rel err", ((out.float() - A.float() @ B.float().t()).norm() / (A.float() @ B.float().t()).norm()).item())
