"""GEMM micro-benchmark (GPU box): the DSFVT GEMM shapes, CUDA-event timed, TFLOP/s."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lvt_b200 import ops
from lvt_b200.ops import Operand

M = int(os.environ.get("GEMM_M", 16384))
reps = int(os.environ.get("GEMM_REPS", 20))
which = sys.argv[1:] or ["qkv", "ffn", "proj", "ffn_dgrad", "ffn_wgrad", "qkv_wgrad", "pv", "softmax"]


GRAPH = os.environ.get("GEMM_GRAPH", "1") == "1"  # time `reps` launches inside one CUDA graph (no host launch cost)


def timeit(fn, flops, name):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if GRAPH:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        g = torch.cuda.CUDAGraph()
        with torch.cuda.stream(side):
            with torch.cuda.graph(g):
                for _ in range(reps):
                    fn()
        torch.cuda.current_stream().wait_stream(side)
        g.replay()
        torch.cuda.synchronize()
        e0.record()
        g.replay()
        e1.record()
    else:
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
    torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / reps * 1e-3
    print(f"{name:12s} {t*1e6:9.1f} us  {flops/t/1e12:8.1f} TFLOP/s")


bf = torch.bfloat16
d, H, da = 512, 8, 128
x = torch.randn(M, d, device="cuda").to(bf)
x2 = torch.randn(M, 2 * d, device="cuda").to(bf)
wq = torch.randn(24, d, da, device="cuda").to(bf)
w = torch.randn(d, d, device="cuda").to(bf)
wp = torch.randn(d, 2 * d, device="cuda").to(bf)
o3 = torch.empty(M, 3 * H * da, device="cuda", dtype=bf)
o1 = torch.empty(M, d, device="cuda", dtype=bf)
of = torch.empty(M, d, device="cuda")
gw = torch.zeros(d, d, device="cuda")
gq = torch.zeros(24, d, da, device="cuda")
L = 256
nb = M // L
P = torch.empty(nb, H, L, L, device="cuda", dtype=bf)
banks = [torch.zeros(H, 1, device="cuda"), torch.zeros(H, 31, device="cuda"), torch.zeros(H, 31, device="cuda")]


def qkv_op(buf, which_, mn):
    return Operand(buf.data_ptr() + 2 * which_ * H * da, 3 * H * da, mn_major=mn, cin=da, zdiv=H, s_zlo=da, s_zhi=L * 3 * H * da)


tests = {
    "qkv": (lambda: ops.gemm(M, 3072, d, Operand(x.data_ptr(), d), Operand(wq.data_ptr(), da, mn_major=True, cin=da, s_blk=d * da), Operand(o3.data_ptr(), 3072), out_bf16=o3), 2.0 * M * 3072 * d),
    "ffn": (lambda: ops.gemm(M, d, d, Operand(x.data_ptr(), d), Operand(w.data_ptr(), d), Operand(o1.data_ptr(), d), out_bf16=o1), 2.0 * M * d * d),
    "proj": (lambda: ops.gemm(M, d, 2 * d, Operand(x2.data_ptr(), 2 * d), Operand(wp.data_ptr(), 2 * d), Operand(of.data_ptr(), d), out_f32=of, res=of), 2.0 * M * d * 2 * d),
    "ffn_dgrad": (lambda: ops.gemm(M, d, d, Operand(x.data_ptr(), d), Operand(w.data_ptr(), d, mn_major=True), Operand(of.data_ptr(), d), out_f32=of), 2.0 * M * d * d),
    "ffn_res": (lambda: ops.gemm(M, d, d, Operand(x.data_ptr(), d), Operand(w.data_ptr(), d), Operand(of.data_ptr(), d), out_f32=of, res=of), 2.0 * M * d * d),
    "ffn_mask": (lambda: ops.gemm(M, d, d, Operand(x.data_ptr(), d), Operand(w.data_ptr(), d, mn_major=True), Operand(o1.data_ptr(), d), out_bf16=o1, aux=x, flags=ops.GEMM_MASK), 2.0 * M * d * d),
    "qkv_dgrad": (lambda: ops.gemm(M, d, 3072, Operand(o3.data_ptr(), 3072), Operand(wq.data_ptr(), da, mn_major=False, cin=da, s_blk=d * da), Operand(o1.data_ptr(), d), out_bf16=o1), 2.0 * M * 3072 * d),
    "ffn_wgrad_auto": (lambda: ops.gemm(d, d, M, Operand(x.data_ptr(), d, mn_major=True), Operand(o1.data_ptr(), d, mn_major=True), Operand(gw.data_ptr(), d), out_f32=gw, splits=-1, flags=ops.GEMM_ATOMIC), 2.0 * M * d * d),
    "qkv_wgrad_auto": (lambda: ops.gemm(d, 3072, M, Operand(x.data_ptr(), d, mn_major=True), Operand(o3.data_ptr(), 3072, mn_major=True), Operand(gq.data_ptr(), da, cin=da, s_blk=d * da), out_f32=gq, splits=-1, flags=ops.GEMM_ATOMIC), 2.0 * M * 3072 * d),
    "ffn_wgrad": (lambda: ops.gemm(d, d, M, Operand(x.data_ptr(), d, mn_major=True), Operand(o1.data_ptr(), d, mn_major=True), Operand(gw.data_ptr(), d), out_f32=gw, splits=int(os.environ.get('WGRAD_SPLITS', 18)), flags=ops.GEMM_ATOMIC), 2.0 * M * d * d),
    "qkv_wgrad": (lambda: ops.gemm(d, 3072, M, Operand(x.data_ptr(), d, mn_major=True), Operand(o3.data_ptr(), 3072, mn_major=True), Operand(gq.data_ptr(), da, cin=da, s_blk=d * da), out_f32=gq, splits=6, flags=ops.GEMM_ATOMIC), 2.0 * M * 3072 * d),
    "pv": (lambda: ops.gemm(L, da, L, Operand(P.data_ptr(), L, zdiv=1, s_zhi=L * L), qkv_op(o3, 2, True), Operand(x2.data_ptr(), H * da, cin=da, zdiv=H, s_zlo=da, s_zhi=L * H * da), out_bf16=x2, batch=nb * H), 2.0 * nb * H * L * L * da),
    "softmax": (lambda: ops.gemm(L, L, da, qkv_op(o3, 0, False), qkv_op(o3, 1, False), Operand(P.data_ptr(), L, zdiv=1, s_zhi=L * L), out_bf16=P, batch=nb * H, alpha=0.088, mode=ops.EPI_SOFTMAX, flags=ops.GEMM_CAUSAL, banks=banks, block=(1, 16, 16), heads=H), 2.0 * nb * H * L * L * da),
}
tests["attn_fused"] = (lambda: ops.gemm(L, L, da, qkv_op(o3, 0, False), qkv_op(o3, 1, False), Operand(P.data_ptr(), L, zdiv=1, s_zhi=L * L), out_bf16=P, batch=nb * H, alpha=0.088, mode=ops.EPI_SOFTMAX, flags=ops.GEMM_CAUSAL, banks=banks, block=(1, 16, 16), heads=H, v=qkv_op(o3, 2, True), o2=Operand(x2.data_ptr(), H * da, cin=da, zdiv=H, s_zlo=da, s_zhi=L * H * da), o2_n=da), 4.0 * nb * H * L * L * da)
tests["attn_fused_nop"] = (lambda: ops.gemm(L, L, da, qkv_op(o3, 0, False), qkv_op(o3, 1, False), Operand(P.data_ptr(), L, zdiv=1, s_zhi=L * L), out_bf16=None, batch=nb * H, alpha=0.088, mode=ops.EPI_SOFTMAX, flags=ops.GEMM_CAUSAL, banks=banks, block=(1, 16, 16), heads=H, v=qkv_op(o3, 2, True), o2=Operand(x2.data_ptr(), H * da, cin=da, zdiv=H, s_zlo=da, s_zhi=L * H * da), o2_n=da), 4.0 * nb * H * L * L * da)
dOb = torch.randn(M, H * da, device="cuda").to(bf)
delta = torch.randn(nb, H, L, device="cuda")
dS = torch.empty_like(P)
dqkv = torch.empty_like(o3)
gbanks = [torch.zeros(H, 1, device="cuda"), torch.zeros(H, 31, device="cuda"), torch.zeros(H, 31, device="cuda")]
tests["attn_bwd_fused"] = (lambda: ops.gemm(L, L, da, Operand(dOb.data_ptr(), H * da, cin=da, zdiv=H, s_zlo=da, s_zhi=L * H * da), qkv_op(o3, 2, False), Operand(dS.data_ptr(), L, zdiv=1, s_zhi=L * L), out_bf16=dS, batch=nb * H, mode=ops.EPI_DS, aux=P, delta=delta, alpha=0.088, v=qkv_op(o3, 1, True), o2=qkv_op(dqkv, 0, False), o2_n=da, banks=gbanks, block=(1, 16, 16), heads=H), 4.0 * nb * H * L * L * da)
lse = torch.randn(nb * H, L, device="cuda") + 8.0
for cz, nm in ((False, "attn_bwd_all"), (True, "attn_bwd_all_causal")):
    tests[nm] = ((lambda cz=cz: ops.attn_bwd(o3, dOb, dqkv, lse, delta, banks, gbanks, nb, H, (1, 16, 16), cz, 0.088)),
                 (8.0 if cz else 10.0) * nb * H * L * L * da)
DBG = int(os.environ.get("GEMM_DBG", "0"))
if DBG:
    tests["qkv"] = (lambda: ops.gemm(M, 3072, d, Operand(x.data_ptr(), d), Operand(wq.data_ptr(), da, mn_major=True, cin=da, s_blk=d * da), Operand(o3.data_ptr(), 3072), out_bf16=o3, flags=DBG), 2.0 * M * 3072 * d)
    tests["ffn"] = (lambda: ops.gemm(M, d, d, Operand(x.data_ptr(), d), Operand(w.data_ptr(), d), Operand(o1.data_ptr(), d), out_bf16=o1, flags=DBG), 2.0 * M * d * d)
for k in which:
    timeit(tests[k][0], tests[k][1], k)
