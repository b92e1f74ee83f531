"""Timeline of pair 0 of the GIN layer kernel gin_fused.cu (MMA issuer, one epilogue warp, one gather warp, the producer).
needs the library built with the hooks:  make -C flowgnn_b200/csrc clean && make -C flowgnn_b200/csrc EXTRA=-DFG_TC2_TRACE
usage: python tools/trace_gin.py   (prints per-tile event offsets in ns for the last layer launch; env FLOWGNN_B200_DBG=1|2|4
additionally switches off the in-edge loads / the h' stores / the z conversion: wrong results, for bottleneck elimination)"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from flowgnn_b200.capi import Context, load_library
from flowgnn_b200.dataset import synthetic_molecules
from flowgnn_b200.weights import load_weights
w = load_weights("gin", os.path.join(ROOT, "tests", "golden", "weights", "GIN"))
big = synthetic_molecules(2048, "molhiv", seed=11).tile(41127)
lib = load_library()
lib.flowgnn_b200_debug_trace.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int]
with Context(0) as c:
    c.load_weights("gin", w); c.upload(big)
    for _ in range(3): c.compute("gin")
    ptr = ctypes.c_void_p()
    assert lib.flowgnn_b200_debug_trace(ctypes.byref(ptr), 1) == 0
    c.compute("gin"); c.synchronize()
    lib.flowgnn_b200_debug_trace(ctypes.byref(ptr), 0)
    buf = np.zeros(4096, dtype=np.uint64)
    torch.cuda.synchronize()
    import ctypes as ct
    cudart = ct.CDLL("libcudart.so.12")
    cudart.cudaMemcpy(buf.ctypes.data_as(ct.c_void_p), ptr, buf.nbytes, 2)
cta = buf[2048:].astype(np.int64)
t = buf[:3 * 64 * 8].reshape(3, 64, 8).astype(np.int64)
t0 = t[0, 0, 0]
np.set_printoptions(linewidth=200)
print("events (ns since the MMA thread started tile 0); MMA: wait_A, got_A, G1 issued, got_A2A, got_A2B, G2 issued")
print("epilogue: got_G1A, arrived_A2A, got_G1B, arrived_A2B, got_G2, stored_H ; gather: start, arrived_A, got_stage, got_A_free, producer_issued_TMA")
for it in list(range(0, 6)) + list(range(30, 36)):
    m = t[0, it, :6] - t0; e = t[1, it, :6] - t0; g = t[2, it, :5] - t0
    print(f"tile {it:2d}  MMA {m}  EPI {e}  GATHER {g}")
per = np.diff(t[0, 8:50, 1])
print("MMA got_A period: mean %.0f ns" % per.mean())
for name, a, b in (("A_FULL wait", (0, 0), (0, 1)), ("G1 issue", (0, 1), (0, 2)), ("G1 issued -> got A2A", (0, 2), (0, 3)), ("A2A -> A2B", (0, 3), (0, 4)), ("A2B -> G2 issued", (0, 4), (0, 5)),
                   ("G2 issued -> next A wait", None, None), ("EPI: got G1A -> arrived A2A (convert a)", (1, 0), (1, 1)), ("EPI: arrived A2A -> got G1B", (1, 1), (1, 2)),
                   ("EPI: convert b", (1, 2), (1, 3)), ("EPI: arrived A2B -> got G2", (1, 3), (1, 4)), ("EPI: H store", (1, 4), (1, 5)),
                   ("MMA G1 issued -> EPI got G1A", (0, 2), (1, 0)), ("EPI arrived A2A -> MMA got A2A", (1, 1), (0, 3)), ("MMA G2 issued -> EPI got G2", (0, 5), (1, 4)),
                   ("GATHER tile time (start -> arrived)", (2, 0), (2, 1)), ("GATHER wait for the stage", (2, 0), (2, 2)), ("GATHER compute (stage full -> all chunks in registers)", (2, 2), (2, 3)),
                   ("GATHER barrier + A stores (-> arrived)", (2, 3), (2, 1)), 
                   ("GATHER arrived -> MMA got A", (2, 1), (0, 1)), ("producer TMA issue -> GATHER got stage", (2, 4), (2, 2))):
    if a is None: continue
    d = t[b[0], 8:50, b[1]] - t[a[0], 8:50, a[1]]
    print(f"{name:45s} mean {d.mean():8.0f} ns  min {d.min():6d} max {d.max():6d}")

# cross-tile intervals: GEMM1 of tile i releases the A tile for the gather of tile i + 1; the stage of tile i + 1 is requested when the gather of tile i is done
d = t[2, 10:52, 4] - t[0, 8:50, 2]
print(f"{'MMA G1(i) issued -> producer issues TMA(i+2)':45s} mean {d.mean():8.0f} ns  min {d.min():6d} max {d.max():6d}")
d = t[2, 9:51, 2] - t[2, 8:50, 1]
print(f"{'GATHER(i) arrived -> GATHER(i+1) got its stage':45s} mean {d.mean():8.0f} ns  min {d.min():6d} max {d.max():6d}")
d = t[2, 9:51, 2] - t[2, 9:51, 4]
print(f"{'producer TMA(i+1) issue -> stage(i+1) full':45s} mean {d.mean():8.0f} ns  min {d.min():6d} max {d.max():6d}")
d = t[0, 9:51, 1] - t[0, 8:50, 5]
print(f"{'MMA G2(i) issued -> got A(i+1)':45s} mean {d.mean():8.0f} ns  min {d.min():6d} max {d.max():6d}")

k0 = np.array([t[1, 0, 6], t[1, 0, 7], t[1, 1, 6], t[1, 1, 7]]) - t0
k1 = np.array([t[1, 2, 6], t[1, 2, 7], t[1, 3, 6], t[1, 3, 7]]) - t0
print("CTA 0   : entry %d, prologue done %d, previous grid done %d, loops done %d  (ns relative to the first tile of pair 0)" % tuple(k0))
print("last CTA: entry %d, prologue done %d, previous grid done %d, loops done %d" % tuple(k1))
nt = int(t[0, 0, 7])
print("tiles:", nt, " pair tiles per pair: %.1f" % (nt / 2 / 74))
last = max(i for i in range(64) if t[0, i, 5] > 0)
print("pair 0: last recorded tile %d, G2 issued at %d ns" % (last, t[0, last, 5] - t0))

dur = (cta[256:256 + 148] - cta[:148]) / 1e3
start = (cta[:148] - cta[:148].min()) / 1e3
print("per-CTA loop time (us): min %.1f mean %.1f max %.1f;  loop start spread %.1f us" % (dur.min(), dur.mean(), dur.max(), start.max()))
print("per pair (even CTA):", np.round(dur[0::2], 0).astype(int).tolist())
