"""Cycle attribution inside k_conv_umma (debug counters via McvdOp.dst2).  usage: umma_timing.py B H Cin Cout [ks]"""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mcvd_b200 import lib
from mcvd_b200.lib import McvdOp
B, H, Cin, Cout = [int(v) for v in sys.argv[1:5]]; ks = int(sys.argv[5]) if len(sys.argv) > 5 else 3
nacc = int(sys.argv[6]) if len(sys.argv) > 6 else 0
mode = sys.argv[7] if len(sys.argv) > 7 else "tab"   # tab | plain | res
dev = "cuda:0"
x = torch.randn(B, H, H, Cin, device=dev); w = torch.randn(ks * ks, Cin, Cout, device=dev) / math.sqrt(Cin * ks * ks)
tab = torch.stack([torch.zeros(B, Cin), torch.ones(B, Cin), torch.ones(B, Cin), torch.zeros(B, Cin)], 2).contiguous().to(dev)
bias = torch.zeros(Cout, device=dev); out = torch.zeros(B, H, H, Cout, device=dev)
kb = lib.umma_kblock(Cin, 0); nt = max(d for d in range(16, 257, 16) if Cout % d == 0)
pk = torch.empty(w.numel() * 4, dtype=torch.uint8, device=dev)
lib.load().mcvd_umma_pack_weights(w.data_ptr(), ks * ks, Cin, Cout, nt, kb, pk.data_ptr(), 5, torch.cuda.current_stream().cuda_stream)
pimg = (H + 1) * (H + 1) if ks == 3 else H * H
o = McvdOp(); o.kind, o.B, o.H, o.W, o.C0, o.Cout, o.i0, o.i1, o.i2 = lib.OP_CONV_UMMA, B, H, H, Cin, Cout, ks, nt, nacc
o.f0, o.f1 = 1.0, 2.0 ** -5
o.src0, o.w, o.bias, o.dst = x.data_ptr(), pk.data_ptr(), bias.data_ptr(), out.data_ptr()
res = torch.randn(B, H, H, Cout, device=dev)
if mode == "tab": o.aux1 = tab.data_ptr(); o.flags = lib.F_ACT_IN
if mode == "res": o.aux0 = res.data_ptr()
arr = lib.make_ops([o]); s = torch.cuda.current_stream().cuda_stream
for _ in range(3): lib.run_program(arr, 1, s)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); [lib.run_program(arr, 1, s) for _ in range(10)]; e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 100
flops = 2.0 * B * H * H * Cin * Cout * ks * ks
print(f"conv {Cin}->{Cout} k{ks} @{H}x{H} B={B} nt={nt} nacc={nacc} kb={kb}: {us:.1f} us/launch, {flops/us/1e6:.1f} TF/s algorithmic ({3*flops/us/1e6:.0f} executed)")
dbg = torch.zeros(148 * 16, dtype=torch.int64, device=dev)
arr[0].aux2 = dbg.data_ptr(); lib.run_program(arr, 1, s); torch.cuda.synchronize()
d = dbg.view(148, 16).double()
d = d[d[:, 0] > 0].mean(0).tolist()
if ks == 1 and os.environ.get("MCVD_CONV1X1", "1") != "0" and Cin % 32 == 0 and Cin <= 320:   # conv1x1_umma.cu slots
    e0.record(); [lib.run_program(arr, 1, s) for _ in range(5)]; e1.record(); torch.cuda.synchronize()
    print(f"  [stationary kernel, DBG build: {e0.elapsed_time(e1) * 200:.1f} us/launch, dbgf={os.environ.get('MCVD_K1_DBGF', '0')}]")
    print(f"  per CTA cycles {d[0]:.0f} | producer t0: wait RAW_FULL {d[1]:.0f} work {d[2]:.0f} | loader: wait B_EMPTY {d[11]:.0f}")
    print(f"  MMA: wait ACC_EMPTY {d[3]:.0f} A_FULL {d[4]:.0f} B_FULL {d[5]:.0f} issue {d[6]:.0f} | epilogue w0: wait ACC_FULL {d[7]:.0f} "
          f"tmem {d[8]:.0f} transpose {d[9]:.0f} store {d[10]:.0f}")
    sys.exit(0)
print(f"  per CTA cycles {d[0]:.0f} | producer t0: wait A_EMPTY {d[1]:.0f} bar {d[2]:.0f} emit {d[3]:.0f} fence+arrive {d[4]:.0f}")
print(f"  MMA: wait ACC_EMPTY {d[5]:.0f} A_FULL {d[6]:.0f} B_FULL {d[7]:.0f} issue {d[8]:.0f} | epilogue w10: wait ACC_FULL {d[9]:.0f} drain {d[10]:.0f}")
