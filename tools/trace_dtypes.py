"""Debug helper: print the dtype flowing through every module of the segmentor under bf16 autocast on cuda:0."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import refign_b200 as P

def desc(t):
    if torch.is_tensor(t):
        return str(t.dtype).replace('torch.', '') + ('/cl' if t.dim() == 4 and t.is_contiguous(memory_format=torch.channels_last) and not t.is_contiguous() else '')
    if isinstance(t, (list, tuple)):
        return '[' + ','.join(desc(u) for u in t) + ']'
    return str(type(t).__name__)

def hook(name):
    def f(m, i, o):
        print("%-50s %-14s %s -> %s" % (name, type(m).__name__, desc(i[0] if len(i) else None), desc(o)))
    return f

torch.manual_seed(0)
bb = P.MixVisionTransformer('mit_b0').cuda()
hd = P.DAFormerHead([32, 64, 160, 256], [0, 1, 2, 3], 19, 'multiple_select').cuda()
for n, m in list(hd.named_modules()) + [("bb." + n, m) for n, m in bb.named_modules() if n.startswith('block1.0') or n.startswith('patch_embed1') or n == 'norm1']:
    if n:
        m.register_forward_hook(hook(n))
x = torch.randn(2, 3, 128, 128, device='cuda')
with torch.autocast('cuda', dtype=torch.bfloat16):
    f = bb(x)
    print([desc(t) for t in f])
    y = hd(f)
