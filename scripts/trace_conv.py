"""Pipeline time stamps of CTA 0 of the sparse-conv kernel (EFGB_LIB_VARIANT=trace build, see spconv_tc.cu EFGB_TC_TRACE).
usage: EFGB_LIB_VARIANT=trace python scripts/trace_conv.py [level ...]"""
import ctypes, os, sys
os.environ.setdefault("EFGB_LIB_VARIANT", "trace")
os.environ["LEVELS"] = ""          # bench_conv builds the geometry; run nothing there
sys.argv = sys.argv[:1] + [a for a in sys.argv[1:]]
lv = [int(a) for a in sys.argv[1:]] or [0, 1]
sys.argv = sys.argv[:1]
sys.path.insert(0, os.path.dirname(__file__))
import numpy as np, torch
import bench_conv as bc
from efg_b200 import ops, _lib

L = _lib.lib()
raw = ctypes.CDLL(_lib.LIB_PATH)
raw.efgb_debug_trace_read.argtypes = [ctypes.c_void_p, ctypes.c_int]
ops.CONV_PRECISION = "bf16x3"
def report(buf, title):
    n = buf.shape[0]
    t0 = buf[0, 0]
    used = int((buf[:, 5] > 0).sum())
    b = buf[:used] - t0
    print(title + ": %d stages traced on CTA 0" % used)
    print("kernel entry %d, setup done %d, epilogue(i): %s, all warps done %d" % (
        b[0, 6], b[1, 6], " ".join("[%d..%d]" % (b[2 + 2 * i, 6], b[3 + 2 * i, 6]) for i in range(4) if buf[3 + 2 * i, 6] > 0), b[0, 7]))
    for s in range(min(used, int(os.environ.get('ROWS', '12')))):
        print("%5d P: wait %d..%d stored/issued %d arrived %d | M: a_full %d issued %d" % (s, b[s, 0], b[s, 1], b[s, 2], b[s, 10], b[s, 4], b[s, 5]))
    e = buf[1000:1016] - t0
    if buf[1000, 3] > 0:
        print("epilogue blocks of the second tile (warp 0): " + " ".join("[ld %d sts %d stg %d | gap %d]" % (
            e[k, 1] - e[k, 0], e[k, 2] - e[k, 1], e[k, 3] - e[k, 2], (e[k + 1, 0] - e[k, 3]) if buf[1001 + k, 0] > 0 else 0) for k in range(16) if buf[1000 + k, 3] > 0))
    if used > 24:
        mm = b[12:used - 4]
        print("MMA warp: a_full -> fence done %.0f, fence -> MMAs issued %.0f, -> commit + syncwarp %.0f, issued(n) -> a_full(n+1) %.0f; period %.0f" % (
            (mm[:, 6] - mm[:, 4]).mean(), (mm[:, 7] - mm[:, 6]).mean(), (mm[:, 5] - mm[:, 7]).mean(), (mm[1:, 4] - mm[:-1, 5]).mean(),
            (mm[-1, 4] - mm[0, 4]) / (len(mm) - 1)))
        print("producer warp 0: a_empty wait %.0f, stores/copies %.0f, -> arrived %.0f, arrived(n) -> wait_begin(n+1) %.0f" % (
            (mm[:, 1] - mm[:, 0]).mean(), (mm[:, 2] - mm[:, 1]).mean(), (mm[:, 10] - mm[:, 2]).mean(), (mm[1:, 0] - mm[:-1, 10]).mean()))


if os.environ.get("DENSE"):
    m = 70688
    for cin, cout in [(256, 256), (256, 1024), (1024, 256)]:
        x = torch.randn(m, cin, device=bc.dev)
        w = torch.randn(cout, 1, cin, device=bc.dev) * 0.05
        for _ in range(3):
            ops.spconv_tc(x, w, None, None, 0)
        torch.cuda.synchronize()
        buf = np.zeros((2048, 12), dtype=np.int64)
        raw.efgb_debug_trace_read(buf.ctypes.data, 2048)
        report(buf, "dense %d -> %d, %d rows" % (cin, cout, m))
    sys.exit(0)

for lvl in lv:
    c = [16, 64, 128, 256][lvl]
    oc, od, nbr, nbr_s, nbr_t, m_in = bc.levels[lvl]
    mo = oc.shape[0]
    feats = torch.randn(mo, c, device=bc.dev)
    w = torch.randn(c, 27, c, device=bc.dev) * 0.05
    packed = torch.empty(L.efgb_spconv_tc_packed_bytes(27, c, c, 2) // 4, dtype=torch.float32, device=bc.dev)
    L.efgb_spconv_tc_pack(ops._p(w), c, 27, c, 0, 2, ops._p(packed), ops._stream())
    out = torch.empty(mo, c, device=bc.dev)
    planes = torch.empty_like(feats)
    L.efgb_split_bf16(ops._p(feats), mo, c, ops._p(planes), ops._stream())
    for _ in range(3):
        L.efgb_spconv_tc_forward_planes(ops._p(planes), mo, c, ops._p(packed), None, ops._p(nbr), mo, 27, c, 0, ops._p(out), ops._stream())
    torch.cuda.synchronize()
    n = 2048
    buf = np.zeros((n, 12), dtype=np.int64)
    raw.efgb_debug_trace_read(buf.ctypes.data, n)
    t0 = buf[0, 0]
    used = int((buf[:, 5] > 0).sum())
    b = buf[:used] - t0
    print("level %d C=%d rows=%d: %d stages traced on CTA 0" % (lvl + 1, c, mo, used))
    print("stage   P:wait_begin  P:wait_end  P:issued   P:done(n)   M:a_full   M:issued   | dP(iter) dM(a_full) wait_empty")
    for s in range(min(used, int(os.environ.get('ROWS', '12')))):
        dp = b[s, 2] - b[s - 1, 2] if s else 0
        dm = b[s, 4] - b[s - 1, 4] if s else 0
        print("%5d %12d %11d %10d %11d %10d %10d   | %7d %9d %9d" % (s, b[s, 0], b[s, 1], b[s, 2], b[s, 3], b[s, 4], b[s, 5], dp, dm, b[s, 1] - b[s, 0]))
    print("kernel entry %d, setup done %d, epilogue(i): %s, all warps done %d" % (
        b[0, 6], b[1, 6], " ".join("[%d..%d]" % (b[2 + 2 * i, 6], b[3 + 2 * i, 6]) for i in range(4) if buf[3 + 2 * i, 6] > 0), b[0, 7]))
    mm = b[12:used - 4]
    print("MMA warp, stages 12..: a_full -> fence done %.0f, fence -> MMAs issued %.0f, MMAs issued -> commit + syncwarp %.0f, issued(n) -> a_full(n+1) %.0f" % (
        (mm[:, 6] - mm[:, 4]).mean(), (mm[:, 7] - mm[:, 6]).mean(), (mm[:, 5] - mm[:, 7]).mean(), (mm[1:, 4] - mm[:-1, 5]).mean()))
    pp = b[12:used - 8]
    # stage n completes in the iteration that issues stage n + 3
    print("producer warp 0, per iteration: a_empty wait %.0f, copies issued %.0f, idx loads + commit %.0f, wait_group %.0f, syncwarp + arrive %.0f, loop back %.0f" % (
        (pp[:, 1] - pp[:, 0]).mean(), (pp[:, 2] - pp[:, 1]).mean(), (b[9:used - 11, 9] - pp[:, 2]).mean(), (b[9:used - 11, 3] - b[9:used - 11, 9]).mean(),
        (b[9:used - 11, 10] - b[9:used - 11, 3]).mean(), (b[13:used - 7, 0] - b[9:used - 11, 10]).mean()))
    mid = b[8:used - 4]
    if len(mid) > 4:
        print("steady state (stages 8..%d): cycles/stage %.0f; producer: wait_empty %.0f, issue %.0f; MMA: a_full->issued %.0f; a_full(n) - P:done(n) %.0f; P:wait_end(n) - M:issued(n - sa) n/a" % (
            used - 4, (mid[-1, 4] - mid[0, 4]) / (len(mid) - 1), (mid[:, 1] - mid[:, 0]).mean(), (mid[:, 2] - mid[:, 1]).mean(),
            (mid[:, 5] - mid[:, 4]).mean(), (mid[:, 4] - mid[:, 3]).mean()))
