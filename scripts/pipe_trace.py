"""clock64 stamps per role of CTA 0 of node_mlp_kernel_pipe64 (gnnseg_mlp_pipe.cu).
    make -C gnn_fpga_b200/csrc trace
    GNNSEG_LIB=gnn_fpga_b200/libgnnseg_trace.so python scripts/pipe_trace.py [n_cols | 0] [hidden_dim]
(the instrumented library is a separate build, see the Makefile; the shipped one carries no stamps)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gnn_fpga_b200 import _lib, SegmentClassifier
dev = torch.device("cuda:0")
h = int(sys.argv[2]) if len(sys.argv) > 2 else 64
n = 100000 if h == 64 else 256000
torch.manual_seed(0)
model = SegmentClassifier(3, h, 8).to(dev).eval()
L = _lib.lib()
L.gnnseg_debug_read_pipe_trace.restype = C.c_int
blob = model.pack_weights()
X4 = torch.randn(n, 4, device=dev)
rows = torch.tanh(torch.randn(n + 1, 5 * h, device=dev))
status = torch.zeros(1, dtype=torch.int32, device=dev)
ptr = lambda t: C.c_void_p(t.data_ptr())
n_cols = int(sys.argv[1]) if len(sys.argv) > 1 and int(sys.argv[1]) > 0 else 5 * h
for _ in range(3):      # h1 in the first h floats of the rows it becomes, as in the forward
    _lib.check(L.gnnseg_state_mlp_step(ptr(blob), ptr(X4), ptr(rows), 5 * h, n, h, ptr(rows), n_cols, ptr(status), None), "state_mlp_step")
torch.cuda.synchronize()
buf = (C.c_longlong * (5 * 16 * 16))()
assert L.gnnseg_debug_read_pipe_trace(buf) == 0
t = np.array(buf, dtype=np.int64).reshape(5, 16, 16)
mma, epi, sto, ldr, wp = t
t0 = mma[0, 0]
r = lambda v: int(v - t0) if v else -1
print("CTA 0, cycles since the MMA warp entered its loop; n_cols = %d" % n_cols)
for it in range(14):
    if mma[it, 0] == 0:
        break
    print("tile %d" % it)
    print("  loader : top %d  A2_EMPTY ok %d  A2 written %d" % (r(ldr[it, 0]), r(ldr[it, 1]), r(ldr[it, 2])))
    print("  mma    : top %d  A2_FULL ok %d  GEMM2 issued %d  A3_FULL ok %d | c0: D3_EMPTY ok %d  WP_FULL ok %d  issued %d | c1: D3_EMPTY ok %d  WP_FULL ok %d  issued %d"
          % tuple(r(mma[it, k]) for k in (0, 1, 2, 3, 4, 5, 6, 8, 9, 10)))
    print("  epi    : top %d  D2_FULL ok %d  H' in registers %d  A3_EMPTY ok %d  A3 written %d" % tuple(r(epi[it, k]) for k in (0, 1, 4, 2, 3)))
    print("  store  : c0: top %d  D3_FULL ok %d  done %d | c1: top %d  D3_FULL ok %d  done %d" % tuple(r(sto[it, k]) for k in range(6)))
print("weight warp loads: (wait start, WP_EMPTY ok, issued)")
for l in range(12):
    if wp[l, 0] == 0:
        break
    print("  load %d: %d %d %d" % (l, r(wp[l, 0]), r(wp[l, 1]), r(wp[l, 2])))
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20):
    _lib.check(L.gnnseg_state_mlp_step(ptr(blob), ptr(X4), ptr(rows), 5 * h, n, h, ptr(rows), n_cols, ptr(status), None), "state_mlp_step")
b.record(); torch.cuda.synchronize()
print("back-to-back launches (warm L2): %.2f us each" % (a.elapsed_time(b) * 1e3 / 20))
cb = (C.c_ulonglong * 320)()
L.gnnseg_debug_read_pipe_cta.restype = C.c_int
assert L.gnnseg_debug_read_pipe_cta(cb) == 0
ct = np.array(cb, dtype=np.uint64).reshape(160, 2)[:148].astype(np.int64)
t0 = ct[:, 0].min()
st, en = (ct[:, 0] - t0) / 1e3, (ct[:, 1] - t0) / 1e3
print("last launch, %%globaltimer: CTA starts %.2f .. %.2f us (median %.2f), ends %.2f .. %.2f us (median %.2f), lifetime %.2f .. %.2f us (median %.2f)"
      % (st.min(), st.max(), np.median(st), en.min(), en.max(), np.median(en), (en - st).min(), (en - st).max(), np.median(en - st)))
order = np.argsort(en)
print("latest CTAs:", [(int(i), round(float(st[i]), 2), round(float(en[i]), 2)) for i in order[-5:]], " earliest:", [(int(i), round(float(en[i]), 2)) for i in order[:3]])
