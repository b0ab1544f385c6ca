#!/usr/bin/env python
"""The other BASELINE.json configs on a GPU box (config 2 is bench.py's line).

    python scripts/run_configs.py [--quick] > gpurun_out/configs.json

cfg1  oavif -t 80 -s 9 --max-pass 6 on a 1024x1024 procedural image: the C++ harness with the CUDA scorer,
      next to the same harness with the CPU oracle injected (checker): same quantizer, same passes,
      byte-identical AVIF; per-stage milliseconds.
cfg3  one 24 MP source, 16 candidates scored in one device pass (device-resident inputs).
cfg4  7680x4320 RGBA, full 6-pass search (alpha encoded, never scored).
cfg5  corpus sweep through the harness's driver (procedural 1920x1080 images, measure.py CSV).
The bundled libaom cannot encode 10-bit, so every real encode below runs with --tenbit 0 (8-bit AV1);
the scorer itself is exercised on 10-bit planes by bench.py and the parity tests.
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oavif_b200.host import harness as H, ssimu2, synth  # noqa: E402

QUICK = "--quick" in sys.argv
out = {"cores": os.cpu_count(), "libavif": H.find_libavif()}


def cfg1():
    from oracle import oracle as O  # checker only
    from test_host_search import OracleScorer
    src = synth.synth(1024, 1024, "mixture", 0)
    o = H.default_opts(tenbit=0, score_tgt=80.0, speed=9, max_pass=6)
    t0 = time.time()
    g, gb = H.search_image(src, o, device=0)
    tg = time.time() - t0
    t0 = time.time()
    c, cb = H.search_image(src, o, scorer=OracleScorer(O))
    tc = time.time() - t0
    b4, bb4 = H.search_image(src, o, batch_width=4, device=0)
    return {
        "gpu": {"q": g.q, "score": g.score, "passes": g.num_pass, "bytes": g.size, "history": g.history(),
                "encode_ms": g.encode_ms, "decode_ms": g.decode_ms, "score_ms": g.score_ms, "total_ms": g.total_ms,
                "wall_s": tg, "log": g.log.decode()},
        "cpu_oracle_scored": {"q": c.q, "score": c.score, "passes": c.num_pass, "bytes": c.size, "history": c.history(),
                              "encode_ms": c.encode_ms, "decode_ms": c.decode_ms, "score_ms": c.score_ms,
                              "total_ms": c.total_ms, "wall_s": tc},
        "same_quantizer": g.q == c.q, "same_passes": g.history() and [q for q, _ in g.history()] == [q for q, _ in c.history()],
        "byte_identical_avif": gb == cb,
        "max_abs_score_delta": max(abs(a[1] - b[1]) for a, b in zip(g.history(), c.history())),
        "batched4": {"q": b4.q, "passes": b4.num_pass, "device_passes": b4.device_passes, "probes": b4.probes,
                     "wasted": b4.wasted, "byte_identical": bb4 == gb, "total_ms": b4.total_ms,
                     "encode_ms": b4.encode_ms, "score_ms": b4.score_ms},
    }


def cfg3():
    import torch
    w, h, n = (6000, 4000, 16) if not QUICK else (3000, 2000, 8)
    src = synth.synth(w, h, "mixture", 3)
    res = {}
    for blur, name in ((ssimu2.BLUR_RECURSIVE, "recursive"), (ssimu2.BLUR_FIR, "fir")):
        with ssimu2.Scorer(w, h, n, blur=blur) as sc:
            sc.set_source(src)
            # candidates: synthetic degradations of increasing strength, carried as 10-bit planes on the device
            dev = []
            bases = [synth.distort(src, 0.1, seed=1), synth.distort(src, 0.5, seed=2)]
            for i in range(n):   # shifted copies of two degradations: distinct buffers, same content class
                d = np.ascontiguousarray(np.roll(bases[i % 2], 3 * i, axis=0))
                y, u, v = synth.rgb8_to_yuv444(d, 10)
                dev.append(tuple(torch.from_numpy(p.view(np.int16)).cuda() for p in (y, u, v)))
            torch.cuda.synchronize()
            ptrs = [[p.data_ptr() for p in cand] for cand in dev]
            sc.score_batch_dev("yuv444", ptrs, [2 * w] * 3, depth=10)
            times = []
            for _ in range(3 if QUICK else 5):
                t0 = time.perf_counter()
                scores = sc.score_batch_dev("yuv444", ptrs, [2 * w] * 3, depth=10)
                times.append(time.perf_counter() - t0)
            tm = sc.timing()
            single = sc.score_batch_dev("yuv444", ptrs[:1], [2 * w] * 3, depth=10)[0]
            res[name] = {"candidates": n, "size": [w, h], "wall_ms_best": min(times) * 1e3, "device_ms": tm.total_ms,
                         "pyramid_ms": tm.pyramid_ms, "blur_ms": tm.blur_ms, "finalize_ms": tm.finalize_ms,
                         "Mpx_per_s": n * w * h / 1e6 / (tm.total_ms / 1e3),
                         "batch_equals_single": single == scores[0], "scores": scores[:4]}
            del dev
            torch.cuda.empty_cache()
    return res


def cfg4():
    w, h = (7680, 4320) if not QUICK else (1920, 1080)
    img = synth.synth_rgba(w, h, "mixture", 4)
    o = H.default_opts(tenbit=0, score_tgt=80.0, speed=9, max_pass=6, auto_tiling=1, quality_alpha=99)
    t0 = time.time()
    r, b = H.search_image(img, o, device=0, want_bytes=False)
    return {"size": [w, h], "q": r.q, "score": r.score, "passes": r.num_pass, "bytes": r.size, "history": r.history(),
            "encode_ms": r.encode_ms, "decode_ms": r.decode_ms, "score_ms": r.score_ms, "total_ms": r.total_ms,
            "wall_s": time.time() - t0, "quality_alpha": 99,
            "note": "reference parser accepts --quality-alpha 0..99 (parse_args.zig:88); 99 used"}


def cfg5():
    import torch
    n_gpu = torch.cuda.device_count()
    count = (96 * n_gpu) if not QUICK else 8
    res = {"n_gpus_visible": n_gpu, "images": count, "size": [1920, 1080]}
    o = H.default_opts(tenbit=0, speed=9, max_pass=6)
    cores = os.cpu_count() or 1
    H.corpus_synth(2, 256, 192, n_gpus=n_gpu, workers_per_gpu=1, opts=o)      # CUDA / libavif warm-up, untimed
    # every GPU brings the same share of host cores for its libaom encodes: cores / visible GPUs each
    wpg = max(1, cores // max(n_gpu, 1))
    res["host_cores_per_gpu"] = wpg
    for g in sorted({1, n_gpu}):
        csvp = os.path.join(ROOT, "gpurun_out", f"corpus_{g}gpu.csv")
        r = H.corpus_synth(count, 1920, 1080, n_gpus=g, workers_per_gpu=wpg, opts=o, csv_path=csvp)
        res[f"{g}gpu"] = {"workers_per_gpu": wpg, "ok": r["ok"], "wall_s": r["wall_s"], "images_per_s": r["ok"] / r["wall_s"],
                          "summary": r["summary"]}
    return res


for name, fn in (("cfg1", cfg1), ("cfg3", cfg3), ("cfg4", cfg4), ("cfg5", cfg5)):
    if any(a.startswith("--only=") for a in sys.argv) and f"--only={name}" not in sys.argv:
        continue
    t0 = time.time()
    try:
        out[name] = fn()
    except Exception as e:  # keep going: one config must not hide the others
        out[name] = {"error": repr(e)}
    out[name]["elapsed_s"] = round(time.time() - t0, 2)
    print(f"[{name}] done in {out[name]['elapsed_s']} s", file=sys.stderr, flush=True)

print(json.dumps(out, indent=1, default=float))
