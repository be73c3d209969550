import sys, os, torch
sys.path.insert(0, "/root/repo")
from isa_b200.attention import scaled_dot_product_attention
dev = torch.device("cuda:0")
torch.manual_seed(0)
q, k, v = [torch.randn(32, 4096, 12, device=dev, requires_grad=True) for _ in range(3)]
for _ in range(2):
    o, _ = scaled_dot_product_attention(q, k, v, 12 ** 0.5)
    o.backward(torch.ones_like(o))
torch.cuda.synchronize()
print("attn", float(o.sum()))
