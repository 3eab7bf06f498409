"""Scratch diagnostic: dumps the tf32 scores of tile 0 / group 0 (lvt_dbg_vq_scores).  Only the round-1 kernel writes them:
run with LVT_VQ_TC1=1."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lvt_b200 import ops, _lib
from oracle import vq as ovq
lib = _lib.require_device()
g = torch.Generator().manual_seed(132)
z = torch.randn((2, 256, 16, 16), generator=g) * 0.3
cb = torch.randn((4, 512, 64), generator=g) * 0.3
want_idx = ovq.vq_argmin_c(z, cb)
lib.lvt_dbg_vq_scores.argtypes = [ctypes.c_void_p]
x = z[0, :64].reshape(64, 256)[:, :128].t()          # tile 0: positions 0..127, group 0 -> [128, 64]
c = cb[0]
want = (c ** 2).sum(1)[None, :] - 2 * x @ c.t()
for layout in ("nhwc", "nchw"):
    dbg = torch.zeros(128, 512, device="cuda")
    lib.lvt_dbg_vq_scores(dbg.data_ptr())
    if layout == "nchw":
        idx = ops.vq_argmin(z.cuda(), cb.cuda())
    else:
        zn = z.permute(0, 2, 3, 1).reshape(-1, 256).contiguous().cuda()
        idx = ops.vq_argmin_nhwc(zn, cb.cuda(), 256)[0].view(2, 4, 16, 16)
    torch.cuda.synchronize()
    got = dbg.cpu()
    print(layout, "score err", (got - want).abs().max().item(), "scale", want.abs().max().item(),
          "idx mismatches", (idx.cpu() != want_idx).sum().item())
lib.lvt_dbg_vq_scores(None)
