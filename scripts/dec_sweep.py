"""Stream-decoder stage time (CUDA-event range dec.speck_decode) for 1 / 8 / 64 chunks of the bench
field and cluster sizes R (SPERR_B200_DEC_CLUSTER); device-resident decompress calls."""
import ctypes as C, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import sperr_b200, bench
L = sperr_b200.load()
dev = torch.device("cuda", 0)
prof_on = L.fn("sperr_b200_prof_enable", None, [C.c_int])
prof_dump = L.fn("sperr_b200_prof_dump", C.c_size_t, [C.c_char_p, C.c_size_t])
sizes = [int(a) for a in sys.argv[1].split(",")] if len(sys.argv) > 1 else [256, 512, 1024]
Rs = sys.argv[2].split(",") if len(sys.argv) > 2 else ["1", "2", "4", "8", "auto"]
for n in sizes:
    dims = (n, n, n)
    vol = bench.field_torch(dims, (0, 0, 0), dev)
    rc, stream = L.compress_3d_dev(vol.data_ptr(), True, dims, (256,) * 3, 3, 1e-3)
    assert rc == 0
    out = torch.empty_like(vol)
    ref = None
    for R in Rs:
        if R == "auto":
            os.environ.pop("SPERR_B200_DEC_CLUSTER", None)
        else:
            os.environ["SPERR_B200_DEC_CLUSTER"] = R
        if R == "1":
            os.environ["SPERR_B200_DECPROF"] = "1"
        else:
            os.environ.pop("SPERR_B200_DECPROF", None)
        for it in range(2):
            rc, d = L.decompress_3d_dev(stream, 0, out.data_ptr(), True)
            assert rc == 0
        os.environ.pop("SPERR_B200_DECPROF", None)
        if ref is None:
            ref = out.clone()
        else:
            assert torch.equal(ref.view(torch.int32), out.view(torch.int32)), "decoded bits depend on R"
        prof_on(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 3
        e0.record()
        for it in range(reps):
            L.decompress_3d_dev(stream, 0, out.data_ptr(), True)
        e1.record()
        torch.cuda.synchronize()
        buf = C.create_string_buffer(1 << 16)
        prof_dump(buf, len(buf))
        prof_on(0)
        st = json.loads(buf.value.decode())
        print("n=%d chunks=%d R=%s decompress %.2f ms  speck_decode %.2f  reconstruct %.2f  outliers %.2f  idwt %.2f" % (
            n, (n // 256) ** 3, R, e0.elapsed_time(e1) / reps, st["dec.speck_decode"]["ms"] / reps,
            st.get("d.reconstruct", {"ms": 0})["ms"] / reps, st.get("d.outliers", {"ms": 0})["ms"] / reps,
            st.get("d.idwt", {"ms": 0})["ms"] / reps), flush=True)
