"""BASELINE.json configs #3, #4, #5 and the noisy variant of #2 on ONE GPU, device resident, each as
one JSON line (same keys as bench.py's line where they apply). The 64 GiB configurations are
measured on the share ONE of eight GPUs gets when the volume is sharded along chunk boundaries
(chunks are independent) -- the line says which box that is.

    python scripts/bench_configs.py [3] [4] [5] [2n] [--share 8] [--reps 3] [--cpu 1]

Roofline: the STAGED-model constants of SURVEY.md 8(d) (bytes per value a one-round-trip-per-stage
pipeline must move), against the measured HBM copy bandwidth; CPU baseline: the unmodified
reference (oracle/_ref) with all host threads on a stated sub-sample."""
import argparse, ctypes as C, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import sperr_b200, bench

GB = 1e9
CK = 256


def field(dims, dtype, dev, noise=0.0):
    v = bench.field_torch(dims, (0, 0, 0), dev)
    if dtype == torch.float64:
        v = v.double()   # the same values, widened (the reference reads f64 input as is)
    if noise > 0.0:
        g = torch.Generator(device=dev); g.manual_seed(99)
        v = v + noise * torch.randn(v.shape, device=dev, dtype=v.dtype, generator=g)
    return v


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        r = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, r


def cpu_ref(vol_np, dims, is_float, mode, q, decomp=True):
    lib, kind, prefix = bench.load_cpu_lib()
    sz, vp = C.c_size_t, C.c_void_p
    comp = getattr(lib, prefix + "comp_3d"); comp.restype = C.c_int
    comp.argtypes = [vp, C.c_int] + [sz] * 6 + [C.c_int, C.c_double, sz, C.POINTER(vp), C.POINTER(sz)]
    dec = getattr(lib, prefix + "decomp_3d"); dec.restype = C.c_int
    dec.argtypes = [vp, sz, C.c_int, sz, C.POINTER(sz), C.POINTER(sz), C.POINTER(sz), C.POINTER(vp)]
    libc = C.CDLL(None); libc.free.argtypes = [vp]
    nt = bench.host_cores()
    dst, n = vp(None), sz(0)
    t0 = time.perf_counter()
    rc = comp(vol_np.ctypes.data_as(vp), int(is_float), *dims, CK, CK, CK, mode, q, nt, C.byref(dst), C.byref(n))
    t1 = time.perf_counter()
    assert rc == 0
    dx, dy, dz, out = sz(0), sz(0), sz(0), vp(None)
    rc = dec(dst, n.value, int(is_float), nt, C.byref(dx), C.byref(dy), C.byref(dz), C.byref(out))
    t2 = time.perf_counter()
    assert rc == 0
    stream = np.ctypeslib.as_array(C.cast(dst, C.POINTER(C.c_uint8)), shape=(n.value,)).copy()
    libc.free(out); libc.free(dst)
    return t1 - t0, t2 - t1, stream, kind, nt


def line(name, workload, dtype, nvals, esz, ms_c, ms_d, bpv_c, bpv_d, stream_bytes, extra):
    peak, src = bench.measured_peak()
    nbytes = nvals * esz
    out = {"config": {"workload": workload}, "metric": "input GB/s (device resident)", "unit": "GB/s", "n_gpus": 1,
           "dtype": dtype, "data": "synthetic", "stream_bytes": int(stream_bytes), "bpp": stream_bytes * 8.0 / nvals}
    if ms_c:
        out["compress_gbs"] = nbytes / (ms_c * 1e-3) / GB
        out["compress_ms"] = ms_c
        ach = bpv_c * nvals / (ms_c * 1e-3) / GB
        out["roofline_compress"] = {"bound": "hbm", "model": "staged, %.1f B/value (SURVEY 8d)" % bpv_c, "achieved": ach,
                                    "peak": peak, "unit": "GB/s", "frac": ach / peak, "peak_source": src}
    if ms_d:
        out["decompress_gbs"] = nbytes / (ms_d * 1e-3) / GB
        out["decompress_ms"] = ms_d
        ach = bpv_d * nvals / (ms_d * 1e-3) / GB
        out["roofline_decompress"] = {"bound": "hbm", "model": "staged, %.1f B/value (SURVEY 8d)" % bpv_d, "achieved": ach,
                                      "peak": peak, "unit": "GB/s", "frac": ach / peak, "peak_source": src}
    out["value"] = out.get("decompress_gbs") if not ms_c else nbytes / ((ms_c + ms_d) * 1e-3) / GB
    out.update(extra)
    print(json.dumps({"bench_config": name, **out}), flush=True)


def cpu_sample(vol, dims_full, is_float, mode, q):
    """reference on the leading cpu_sample_dims() box of the same field"""
    sd, cores = bench.cpu_sample_dims()
    sd = tuple(min(a, b) for a, b in zip(sd, dims_full))
    v3 = vol.view(dims_full[2], dims_full[1], dims_full[0])[:sd[2], :sd[1], :sd[0]].contiguous().cpu().numpy().reshape(-1)
    tc, td, stream, kind, nt = cpu_ref(v3, sd, is_float, mode, q)
    nb = v3.size * v3.itemsize
    return {"cpu_baseline": {"compress_gbs": nb / tc / GB, "decompress_gbs": nb / td / GB, "value": nb / (tc + td) / GB,
                             "unit": "GB/s", "cores": nt, "kind": kind,
                             "sample": "%dx%dx%d leading box of the same field, all host threads" % sd}}, sd, stream


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("which", nargs="*", default=["3", "4", "5", "2n"])
    ap.add_argument("--share", type=int, default=8, help="measure the box one of SHARE GPUs gets (1: the whole volume)")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--cpu", type=int, default=1)
    a = ap.parse_args(argv)
    L = sperr_b200.load()
    dev = torch.device("cuda", 0)

    def chunk_streams_equal(stream, ref_stream, sd, dims):
        """the chunks of the CPU sample box are the leading chunks of the volume only when the box
        spans whole rows of chunks; compare chunk 0 always"""
        from sperr_b200 import sharded
        v, c, isf, hlen, lens = sharded.parse_container(L.lib, stream)
        v2, c2, isf2, hlen2, lens2 = sharded.parse_container(L.lib, ref_stream)
        return bool(np.array_equal(stream[hlen:hlen + int(lens[0])], ref_stream[hlen2:hlen2 + int(lens2[0])]))

    for w in a.which:
        if w == "3":   # 2048^3 f64, fixed rate 2 bpp, 256^3 chunks, sharded over 8 GPUs
            full = (2048, 2048, 2048)
            dims = (2048, 2048, 2048 // a.share) if a.share > 1 else full
            vol = field(dims, torch.float64, dev)
            n = vol.numel()
            ms_c, (rc, stream) = timed(lambda: L.compress_3d_dev(vol.data_ptr(), False, dims, (CK,) * 3, 1, 2.0), a.reps)
            assert rc == 0
            out = torch.empty_like(vol)
            ms_d, (rc, _) = timed(lambda: L.decompress_3d_dev(stream, 0, out.data_ptr(), False), a.reps)
            assert rc == 0
            extra = {}
            if a.cpu:
                extra, sd, rs = cpu_sample(vol, dims, False, 1, 2.0)
                extra["chunk0_equals_cpu_reference"] = chunk_streams_equal(stream, rs, sd, dims)
            line("3", "synthetic %dx%dx%d fp64 (the box 1 of %d GPUs holds of 2048^3), fixed rate 2 bpp, 256^3 chunks (%d)"
                 % (dims + (a.share, n // CK ** 3)), "f64", n, 8, ms_c, ms_d, 44.0, 26.5, stream.size, extra)
            del vol, out
        elif w == "4":   # 4096x4096x1024 f32, PSNR target, decompression only
            full = (4096, 4096, 1024)
            # a box of whole 256^3 chunks: the z axis has only 4 chunk slabs, so 8 shares split z by 4 and y by 2
            zs = min(a.share, 4)
            dims = (4096, 4096 // (a.share // zs), 1024 // zs) if a.share > 1 else full
            vol = field(dims, torch.float32, dev)
            n = vol.numel()
            rc, stream = L.compress_3d_dev(vol.data_ptr(), True, dims, (CK,) * 3, 2, 80.0)
            assert rc == 0
            out = torch.empty_like(vol)
            ms_d, (rc, _) = timed(lambda: L.decompress_3d_dev(stream, 0, out.data_ptr(), True), a.reps)
            assert rc == 0
            rng = float(vol.max() - vol.min())
            sq, step = 0.0, 1 << 28   # in slabs: the fp64 difference of 16 Gi values does not fit beside the volumes
            fo, fv = out.reshape(-1), vol.reshape(-1)
            for a0 in range(0, n, step):
                sq += float(((fo[a0:a0 + step].double() - fv[a0:a0 + step].double()) ** 2).sum())
            mse = sq / n
            extra = {"psnr_db": 10 * np.log10(rng * rng / mse)}
            if a.cpu:
                e2, sd, rs = cpu_sample(vol, dims, True, 2, 80.0)
                extra.update(e2)
            line("4", "synthetic %dx%dx%d fp32 (the box 1 of %d GPUs holds of 4096x4096x1024), PSNR 80 dB, 256^3 chunks "
                 "(%d), decompression only" % (dims + (a.share, n // CK ** 3)), "f64", n, 4, None, ms_d, 36.0, 22.5,
                 stream.size, extra)
            del vol, out
        elif w == "5":   # 4096 slices of 2048^2 f32, PWE, batched slice entry points
            ns_total = 4096 // a.share if a.share > 1 else 4096
            batch = min(256, ns_total)
            vol = field((2048, 2048, batch), torch.float32, dev)   # every batch codes the same 256 slices
            n = vol.numel()
            def comp():
                return L.compress_2d_batch(vol.data_ptr(), True, (2048, 2048), batch, 3, 1e-3, device=True)
            ms_c, (rc, streams, lens) = timed(comp, a.reps)
            assert rc == 0
            out = torch.empty_like(vol)
            ms_d, (rc, _) = timed(lambda: L.decompress_2d_batch(streams, lens, (2048, 2048), True, d_out_ptr=out.data_ptr()), a.reps)
            assert rc == 0
            nb = ns_total // batch
            extra = {"max_abs_err": float((out - vol).abs().max()), "batches": nb,
                     "slices_per_batch": batch}
            line("5", "%d slices of 2048^2 fp32 (1 of %d GPUs' share of 4096), PWE 1e-3, in %d batches of %d slices (rates per batch)"
                 % (ns_total, a.share, nb, batch), "f64", n, 4, ms_c, ms_d, 60.0, 25.0, int(np.sum(lens)), extra)
            del vol, out
        elif w == "2n":   # config #2 with a noise floor at the tolerance
            dims = (1024, 1024, 1024)
            vol = field(dims, torch.float32, dev, noise=1e-3)
            n = vol.numel()
            ms_c, (rc, stream) = timed(lambda: L.compress_3d_dev(vol.data_ptr(), True, dims, (CK,) * 3, 3, 1e-3), a.reps)
            assert rc == 0
            out = torch.empty_like(vol)
            ms_d, (rc, _) = timed(lambda: L.decompress_3d_dev(stream, 0, out.data_ptr(), True), a.reps)
            assert rc == 0
            extra = {"max_abs_err": float((out - vol).abs().max())}
            if a.cpu:
                e2, sd, rs = cpu_sample(vol, dims, True, 3, 1e-3)
                extra.update(e2)
                extra["chunk0_equals_cpu_reference"] = chunk_streams_equal(stream, rs, sd, dims)
            line("2-noisy", "synthetic 1024^3 fp32 + N(0, 1e-3) noise, PWE tol 1e-3, 256^3 chunks (64)", "f64", n, 4, ms_c, ms_d,
                 54.0, 22.5, stream.size, extra)
            del vol, out
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
