"""Measured FP64 peak of this GPU (SURVEY.md 8d C1 / VERDICT r1 missing #3): register-resident DFMA chains on all SMs
(tools/probes/dfma_peak.cu).  Prints one JSON object; bench.py runs the same probe live for its FP64 roofline."""
import ctypes as C, json, os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load():
    L = C.CDLL(os.path.join(R, "tools", "_build", "libfadb_probes.so"))
    L.fadb_probe_dfma_peak.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double)]
    L.fadb_probe_pcie.argtypes = [C.c_size_t, C.c_size_t, C.c_int, C.c_int, C.POINTER(C.c_double)]
    return L


def dfma_peak(L=None, ctas_per_sm=8, chains=8, iters=4096):
    L = L or load()
    out = (C.c_double * 3)()
    rc = L.fadb_probe_dfma_peak(ctas_per_sm, chains, iters, out)
    if rc:
        raise RuntimeError("dfma probe failed: %d" % rc)
    return {"fp64_inst_per_s": out[0], "tflops": 2 * out[0] / 1e12, "ms": out[1], "sms": int(out[2]),
            "ctas_per_sm": ctas_per_sm, "chains": chains}


if __name__ == "__main__":
    L = load()
    rows = [dfma_peak(L, c, ch) for c, ch in ((8, 8), (8, 4), (4, 8), (3, 8), (2, 8), (8, 1), (3, 1), (3, 2))]
    best = max(rows, key=lambda r: r["fp64_inst_per_s"])
    print(json.dumps({"best": best, "sweep": rows}, indent=1))
