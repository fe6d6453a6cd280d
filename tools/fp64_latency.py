"""Dependent-issue latency of the FP64 instructions on this GPU (tools/probes/fp64_latency.cu): cycles per instruction on one
dependent chain of one warp.  Prints one JSON object."""
import ctypes as C, json, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
L = C.CDLL(os.path.join(R, "tools", "_build", "libfadb_probes.so"))
L.fadb_probe_fp64_latency.argtypes = [C.POINTER(C.c_double)]
out = (C.c_double * 5)()
rc = L.fadb_probe_fp64_latency(out)
assert rc == 0, rc
print(json.dumps({"cycles_per_dependent_instruction": {"DFMA": round(out[0], 2), "DMUL": round(out[1], 2), "DADD": round(out[2], 2),
                                                       "MUFU.RSQ64H (rsqrt.approx.ftz.f64)": round(out[3], 2), "SHFL (64-bit)": round(out[4], 2)}}))
