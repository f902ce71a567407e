"""ncu target: the epilogue/latency-bound short-K convs (1x1 skip-conv dgrad 128 -> 128 at 256^2; 3x3 dgrad 128 -> 128)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from stain2stain_b200 import kernels as K  # noqa: E402

B = 64
dev = "cuda"
g = K.from_float(torch.randn(B, 256, 256, 128, device=dev), K.GRAD)
w1 = K.from_float(torch.randn(128, 128, device=dev) * 0.05, K.GRAD)
w9 = K.from_float(torch.randn(128, 9 * 128, device=dev) * 0.02, K.GRAD)
out = torch.empty_like(g)
for _ in range(2):
    K.conv_fwd([(g, 1, 1)], w1, 128, 256, 256, a_fmt=K.GRAD, w_fmt=K.GRAD, out_fmt=K.GRAD, out=out)
    K.conv_fwd([(g, 9, 1)], w9, 128, 256, 256, a_fmt=K.GRAD, w_fmt=K.GRAD, out_fmt=K.GRAD, out=out)
torch.cuda.synchronize()
e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
e[0].record()
K.conv_fwd([(g, 1, 1)], w1, 128, 256, 256, a_fmt=K.GRAD, w_fmt=K.GRAD, out_fmt=K.GRAD, out=out)
e[1].record()
K.conv_fwd([(g, 9, 1)], w9, 128, 256, 256, a_fmt=K.GRAD, w_fmt=K.GRAD, out_fmt=K.GRAD, out=out)
e[2].record()
torch.cuda.synchronize()
print("1x1 128->128@256^2 B64: %.3f ms; 3x3: %.3f ms" % (e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])))
