"""Stream-decoder phase profile (SPERR_B200_DECPROF) on the NOISY variant of the bench field
(N(0, tol) added: the last bit planes dominate), n^3 volume, 256^3 chunks."""
import ctypes as C, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import sperr_b200, bench
L = sperr_b200.load()
dev = torch.device("cuda", 0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
noise = float(sys.argv[2]) if len(sys.argv) > 2 else 1e-3
prof_on = L.fn("sperr_b200_prof_enable", None, [C.c_int])
prof_dump = L.fn("sperr_b200_prof_dump", C.c_size_t, [C.c_char_p, C.c_size_t])
dims = (n, n, n)
vol = bench.field_torch(dims, (0, 0, 0), dev)
g = torch.Generator(device=dev); g.manual_seed(99)
vol = vol + noise * torch.randn(vol.shape, device=dev, dtype=vol.dtype, generator=g)
rc, stream = L.compress_3d_dev(vol.data_ptr(), True, dims, (256,) * 3, 3, 1e-3)
assert rc == 0
print("n=%d noise=%g bpp=%.3f" % (n, noise, stream.size * 8.0 / vol.numel()), flush=True)
out = torch.empty_like(vol)
for R in (["1", "auto"] if n <= 512 else ["auto"]):
    if R == "auto":
        os.environ.pop("SPERR_B200_DEC_CLUSTER", None)
    else:
        os.environ["SPERR_B200_DEC_CLUSTER"] = R
    os.environ["SPERR_B200_DECPROF"] = "1"
    rc, d = L.decompress_3d_dev(stream, 0, out.data_ptr(), True)
    assert rc == 0
    os.environ.pop("SPERR_B200_DECPROF", None)
    prof_on(1)
    for it in range(2):
        L.decompress_3d_dev(stream, 0, out.data_ptr(), True)
    torch.cuda.synchronize()
    buf = C.create_string_buffer(1 << 16)
    prof_dump(buf, len(buf)); prof_on(0)
    st = json.loads(buf.value.decode())
    print("R=%s" % R, " ".join("%s=%.2f" % (k, v["ms"] / 2) for k, v in sorted(st.items())), flush=True)
prof_on(1)
for it in range(2):
    L.compress_3d_dev(vol.data_ptr(), True, dims, (256,) * 3, 3, 1e-3)
torch.cuda.synchronize()
buf = C.create_string_buffer(1 << 16)
prof_dump(buf, len(buf)); prof_on(0)
st = json.loads(buf.value.decode())
print("compress", " ".join("%s=%.2f" % (k, v["ms"] / 2) for k, v in sorted(st.items())), flush=True)
