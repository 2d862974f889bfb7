"""Bring-up / timing tool for the CTA-pair conv kernel (csrc/conv_umma2.cu): one conv through the C-ABI,
checked against a float64 torch evaluation (small shapes) or against the round-1 kernel (big shapes), timed
with CUDA events, with the in-kernel cycle counters.

usage: conv2_check.py B H C0 C1 Cout ks [flags]     flags: t=norm table+SiLU, r=residual, s<C2>[,<C3>]=fused 1x1
                                                    shortcut, g=epilogue GroupNorm statistics, n<NT>=n tile,
                                                    m<k>=split mode, q=skip the CPU reference (compare with v1),
                                                    d<bits>=timing switches (1 no producer work, 2 no epilogue, 4 no weight reloads)
"""
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

from mcvd_b200 import lib
from mcvd_b200.lib import McvdOp

B, H, C0, C1, Cout, ks = [int(v) for v in sys.argv[1:7]]
flags = sys.argv[7:]
use_tab = "t" in flags
use_res = "r" in flags
use_stats = "g" in flags
quick = "q" in flags
C2 = C3 = 0
NT = lib.umma2_pick_nt(Cout, ks)
split = 3
dbgf = 0
for f in flags:
    if f.startswith("s"):
        cs = [int(v) for v in f[1:].split(",")]
        C2, C3 = cs[0], (cs[1] if len(cs) > 1 else 0)
    if f.startswith("n"):
        NT = int(f[1:])
    if f.startswith("m"):
        split = int(f[1:])
    if f.startswith("d"):
        dbgf = int(f[1:])
dev = "cuda:0"
Cin, Cs = C0 + C1, C2 + C3
g = torch.Generator().manual_seed(1)
rn = lambda *s: torch.randn(*s, generator=g)
x0, x1 = rn(B, H, H, C0), (rn(B, H, H, C1) if C1 else None)
y0, y1 = (rn(B, H, H, C2) if C2 else None), (rn(B, H, H, C3) if C3 else None)
w = rn(Cout, Cin, ks, ks) / math.sqrt(Cin * ks * ks)
w2 = rn(Cout, Cs, 1, 1) / math.sqrt(max(Cs, 1)) if Cs else None
bias = rn(Cout) * 0.1
res = rn(B, H, H, Cout) if use_res else None
tab = torch.stack([rn(B, Cin) * 0.3, 0.5 + torch.rand(B, Cin, generator=g), 1 + 0.3 * rn(B, Cin), 0.2 * rn(B, Cin)],
                  2).contiguous() if use_tab else None
scale = 0.7071
taps = lambda ww: ww.permute(2, 3, 1, 0).reshape(ww.shape[2] * ww.shape[3], ww.shape[1], ww.shape[0]).contiguous()
L = lib.load()
stream = torch.cuda.current_stream().cuda_stream
d = lambda t: None if t is None else t.to(dev).contiguous()
x0d, x1d, y0d, y1d, bd, rd, td = d(x0), d(x1), d(y0), d(y1), d(bias), d(res), d(tab)
tab3 = None
if use_tab:
    tab3 = torch.stack([tab[..., 0], tab[..., 1] * tab[..., 2], tab[..., 3]], 1).contiguous().to(dev)
t1 = taps(w).to(dev)
t2 = taps(w2).to(dev) if Cs else None
amax = float(max(t1.abs().max(), t2.abs().max() if Cs else 0))
k = int(math.floor(math.log2(512.0 / amax)))

kb = lib.umma2_plan(H, H, ks, C0, C1, C2, C3, NT, use_stats)
info = lib.umma2_plan_info(H, H, ks, C0, C1, C2, C3, NT, use_stats)
print(f"conv {Cin}(+{Cs} shortcut)->{Cout} k{ks} @{H}x{H} B={B} nt={NT} plan={info}")
assert kb, "no plan"
per_unit = (Cin // kb) * ks * ks + Cs // kb
pk = torch.empty((ks * ks * Cin + Cs) * Cout * 4, dtype=torch.uint8, device=dev)
assert L.mcvd_umma2_pack_weights(t1.data_ptr(), ks * ks, Cin, Cout, NT, kb, pk.data_ptr(), k, 0, per_unit, stream) > 0
if Cs:
    assert L.mcvd_umma2_pack_weights(t2.data_ptr(), 1, Cs, Cout, NT, kb, pk.data_ptr(), k, (Cin // kb) * ks * ks,
                                     per_unit, stream) > 0
out = torch.zeros(B, H, H, Cout, device=dev)
stats = torch.zeros(lib.umma2_stats_bytes(B, H, H, ks, Cout) // 8, dtype=torch.int64, device=dev) if use_stats else None
o = McvdOp()
o.kind, o.B, o.H, o.W, o.C0, o.C1, o.Cout, o.i0, o.i1, o.i2, o.i3 = lib.OP_CONV_UMMA2, B, H, H, C0, C1, Cout, ks, NT, kb, split
o.f0, o.f1 = scale, 2.0 ** -k
o.src0, o.src1, o.w, o.bias, o.dst = x0d.data_ptr(), (x1d.data_ptr() if C1 else None), pk.data_ptr(), bd.data_ptr(), out.data_ptr()
if use_res: o.aux0 = rd.data_ptr()
if use_tab: o.aux1 = tab3.data_ptr(); o.flags = lib.F_ACT_IN
if Cs:
    o.src2, o.C2 = y0d.data_ptr(), C2
    if C3: o.src3, o.C3 = y1d.data_ptr(), C3
if use_stats: o.dst2 = stats.data_ptr()
o.i7 = dbgf
arr = lib.make_ops([o])
lib.run_program(arr, 1, stream)
torch.cuda.synchronize()
print("  ran")

# ---- reference
if not quick:
    xin = x0 if x1 is None else torch.cat([x0, x1], 3)
    if use_tab:
        t = tab.view(B, 1, 1, Cin, 4)
        xin = ((xin - t[..., 0]) * t[..., 1]) * t[..., 2] + t[..., 3]
        xin = xin * torch.sigmoid(xin)
    ref = F.conv2d(xin.permute(0, 3, 1, 2).double(), w.double(), bias.double(), padding=ks // 2).permute(0, 2, 3, 1)
    if Cs:
        ys = y0 if y1 is None else torch.cat([y0, y1], 3)
        ref = ref + F.conv2d(ys.permute(0, 3, 1, 2).double(), w2.double()).permute(0, 2, 3, 1)
    if use_res:
        ref = ref + res.double()
    ref = (ref * scale).float()
    err = (out.cpu() - ref).abs().max().item()
    print(f"  max abs err vs float64 reference {err:.3e} (ref max {ref.abs().max().item():.2f}) "
          f"{'OK' if err < 2e-5 * max(1.0, ref.abs().max().item()) else 'MISMATCH'}")
    if err > 1e-3:
        bad = ((out.cpu() - ref).abs() > 1e-3)
        idx = bad.nonzero()
        print("  mismatching elements:", int(bad.sum()), "of", bad.numel(), "first", idx[:4].tolist(), "last", idx[-2:].tolist())
        print("  bad by channel block of 16:", bad.reshape(-1, Cout // 16, 16).any(2).sum(0).tolist()[:16])
        print("  bad by image:", bad.reshape(B, -1).any(1).tolist())
else:
    # round-1 kernel as the GPU-side reference (no shortcut / same packing needs -> plain comparison only)
    kb1 = lib.umma_kblock(C0, C1)
    if Cs:
        kb1 = min(kb1, lib.umma_kblock(C2, C3))
    parts = []
    for t_ in ([t1, t2] if Cs else [t1]):
        p_ = torch.empty(t_.numel() * 4, dtype=torch.uint8, device=dev)
        assert L.mcvd_umma_pack_weights(t_.data_ptr(), t_.shape[0], t_.shape[1], Cout, NT, kb1, p_.data_ptr(), k, stream) > 0
        parts.append(p_.view(Cout // NT, -1))
    pk1 = torch.cat(parts, 1).contiguous().view(-1)
    out1 = torch.zeros_like(out)
    o1 = McvdOp()
    o1.kind, o1.B, o1.H, o1.W, o1.C0, o1.C1, o1.Cout, o1.i0, o1.i1 = lib.OP_CONV_UMMA, B, H, H, C0, C1, Cout, ks, NT
    o1.f0, o1.f1 = scale, 2.0 ** -k
    o1.src0, o1.src1, o1.w, o1.bias, o1.dst = x0d.data_ptr(), (x1d.data_ptr() if C1 else None), pk1.data_ptr(), bd.data_ptr(), out1.data_ptr()
    if use_res: o1.aux0 = rd.data_ptr()
    if use_tab: o1.aux1 = td.data_ptr(); o1.flags = lib.F_ACT_IN
    if Cs:
        o1.src2, o1.C2 = y0d.data_ptr(), C2
        if C3: o1.src3, o1.C3 = y1d.data_ptr(), C3
    arr1 = lib.make_ops([o1])
    for _ in range(3): lib.run_program(arr1, 1, stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); [lib.run_program(arr1, 1, stream) for _ in range(10)]; e1.record(); torch.cuda.synchronize()
    us1 = e0.elapsed_time(e1) * 100
    err = (out - out1).abs().max().item()
    print(f"  max abs diff vs round-1 kernel {err:.3e} {'OK' if err < 4e-5 * max(1.0, out1.abs().max().item()) else 'MISMATCH'}; "
          f"round-1 kernel {us1:.1f} us")

if use_stats:
    pimg = (H + 1) * (H + 1) if ks == 3 else H * H
    nj, ntiles = 127 // pimg + 2, 2 * ((B * pimg + 255) // 256)
    st = stats.cpu().view(ntiles, nj, 2, Cout)
    y = out.cpu()
    xi = torch.round(y.double() * 65536.0).clamp(-(1 << 28), 1 << 28).to(torch.int64)
    exp = torch.zeros_like(st)
    yy, xx = torch.meshgrid(torch.arange(H), torch.arange(H), indexing="ij")
    r = ((yy + 1) * (H + 1) + xx + 1) if ks == 3 else (yy * H + xx)
    for b in range(B):
        q = (b * pimg + r).reshape(-1)
        tt = q // 128
        jj = b - torch.clamp((tt * 128) // pimg, max=B - 1)
        v = xi[b].reshape(H * H, Cout)
        exp[:, :, 0].index_put_((tt, jj), v, accumulate=True)
        exp[:, :, 1].index_put_((tt, jj), v * v, accumulate=True)
    print("  epilogue statistics exact:", bool(torch.equal(st, exp)), "mismatches", int((st != exp).sum()))

for _ in range(3): lib.run_program(arr, 1, stream)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); [lib.run_program(arr, 1, stream) for _ in range(10)]; e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 100
flops = 2.0 * B * H * H * (Cin * ks * ks + Cs) * Cout
print(f"  {us:.1f} us/launch, {flops / us / 1e6:.1f} TF/s algorithmic ({3 * flops / us / 1e6:.0f} executed)")
dbg = torch.zeros(148 * 16, dtype=torch.int64, device=dev)
arr[0].aux2 = dbg.data_ptr(); lib.run_program(arr, 1, stream); torch.cuda.synchronize()
dd = dbg.view(148, 16).double()
lead = dd[0::2]; lead = lead[lead[:, 0] > 0].mean(0).tolist()
peer = dd[1::2]; peer = peer[peer[:, 0] > 0].mean(0).tolist()
print(f"  leader CTA cycles {lead[0]:.0f} | producer t0: wait A_EMPTY {lead[1]:.0f} transform {lead[3]:.0f} fence+arrive {lead[4]:.0f}")
print(f"  MMA: wait ACC_EMPTY {lead[5]:.0f} A_FULL(+peer) {lead[6]:.0f} B_FULL(+peer) {lead[7]:.0f} issue {lead[8]:.0f} | "
      f"epilogue: wait ACC_FULL {lead[9]:.0f} drain {lead[10]:.0f}")
print(f"  peer CTA cycles {peer[0]:.0f} | producer t0: wait A_EMPTY {peer[1]:.0f} transform {peer[3]:.0f} | epilogue: wait {peer[9]:.0f} drain {peer[10]:.0f}")
