"""Per-phase softmax cycles of the attention kernels (LR_ATTN_TRACE): persistent vs one-CTA-per-item (LR_ATTN_NO_PERSIST=1)."""
import os
import sys

os.environ.setdefault("LR_ATTN_TRACE", "1")
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402
from gpu_trace import show_attn  # noqa: E402
from leftrefill_b200 import ops  # noqa: E402

torch.manual_seed(0)
q = torch.randn(8, 8192, 320, device="cuda").half()
show_attn("attention b=8 h=5 8192x8192", lambda: ops.attention(q, q, q, 5))
q2 = torch.randn(8, 2048, 640, device="cuda").half()
show_attn("attention b=8 h=10 2048x2048", lambda: ops.attention(q2, q2, q2, 10))
c = torch.randn(8, 77, 320, device="cuda").half()
show_attn("attention b=8 h=5 8192x77", lambda: ops.attention(q, c, c, 5))
