"""Join an ncu launch list of one cfg forward with the lowered program: conv_umma time by shape."""
import csv, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from mcvd_b200 import lib
from mcvd_b200.synthetic import make_module
from mcvd_b200.program import Engine
from op_interpreter import Interpreter
path, name, B = sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "cfg2", int(sys.argv[3]) if len(sys.argv) > 3 else 64
lines = [l for l in open(path) if not l.startswith("==")]
times = []
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum": continue
    v = float(row["Metric Value"].replace(",", "")); u = row["Metric Unit"]
    times.append(v / 1e3 if u == "ns" else v if u == "us" else v * 1e3)
cfg, net, sd = make_module(name, "cpu")
eng = Engine(net, _test_backend=Interpreter()); net._engine = eng
ops = eng.program(B).step_ops
assert len(ops) == len(times), (len(ops), len(times))
tot = {}
for op, us in zip(ops, times):
    if op.kind == lib.OP_CONV_UMMA:
        fl = 2.0 * op.B * op.H * op.W * (op.C0 + op.C1) * op.Cout * op.i0 * op.i0
        t = tot.setdefault((op.i0, op.H, op.C0 + op.C1, op.Cout, bool(op.aux1)), [0, 0.0, fl]); t[0] += 1; t[1] += us
s = sum(v[1] for v in tot.values())
k1 = sum(v[1] for k, v in tot.items() if k[0] == 1)
print(f"conv_umma total {s:.0f} us; 1x1 convs {k1:.0f} us; (ks,H,Cin,Cout,fused-norm): n, total us, us each, TF/s algorithmic")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[4]) if len(sys.argv) > 4 else 28]:
    print(" ", k, v[0], "%.0f" % v[1], "%.0f" % (v[1] / v[0]), "%.0f" % (v[2] / (v[1] / v[0]) / 1e6))
