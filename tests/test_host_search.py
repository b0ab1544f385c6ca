"""The C++ host harness around the scored path: libavif glue (src/io.zig), one-image search
(src/main.zig:86-116) and the corpus loop (scripts/measure.py).

CPU part: the harness with the CPU oracle INJECTED as scorer must walk exactly the trace that a
plain Python loop (transcribed policy + libavif + oracle) walks, and write the bytes a direct encode
at the chosen q produces.  GPU part: with the CUDA scorer it must choose the same quantizer, take the
same passes and emit byte-identical AVIF (north_star), sequentially and in batched mode.

The bundled libaom cannot ENCODE 10-bit, so every real encode here runs with --tenbit 0."""
import csv
import os

import numpy as np
import pytest

from oavif_b200.host import harness as H
from oavif_b200.host import synth
from test_tq_policy import z_search

pytestmark = pytest.mark.skipif(H.find_libavif() is None, reason="no libavif in this image")


class OracleScorer:
    def __init__(self, oracle, mode=0):
        self.O, self.mode, self.n = oracle, mode, 0

    def set_source(self, rgb):
        self.src = rgb

    def score(self, y, u, v, depth, matrix, rgba):
        self.n += 1
        return self.O.ssimu2_rgb8(self.src, self.O.yuv444_to_rgb8(y, u, v, depth, matrix, rgba), self.mode)


def opts(**kw):
    return H.default_opts(tenbit=0, **kw)


def test_default_options_are_the_reference_defaults():
    o = H.default_opts()
    assert (o.quality_alpha, o.speed, o.max_threads, o.auto_tiling, o.score_tgt, o.tenbit, o.tune, o.tolerance, o.max_pass,
            o.quality, o.color_primaries, o.transfer_characteristics, o.matrix_coefficients) == \
           (0, 9, 1, 1, 80.0, 1, b"iq", 2.0, 6, -1, 2, 2, 2)          # parse_args.zig:49-63


def test_encode_is_deterministic_and_decode_matches_fixture(golden_dir):
    g = np.load(os.path.join(golden_dir, "avif_roundtrip.npz"))
    src = g["src"]
    a = H.encode(src, 65, opts())
    assert a == H.encode(src, 65, opts()) and len(a) == int(g["avif_bytes65"])
    np.testing.assert_array_equal(H.decode_rgb8(a, src.shape[1], src.shape[0]), g["rgb65"])
    assert len(H.encode(src, 40, opts())) == int(g["avif_bytes40"])


def test_ten_bit_encode_fails_loudly_with_this_libaom():
    src = synth.synth(64, 48, "mixture", 1)
    with pytest.raises(RuntimeError, match="AddImageFailed"):
        H.encode(src, 50, H.default_opts())                          # tenbit = 1 (the reference default)


def test_search_with_injected_oracle_equals_a_plain_python_loop(oracle):
    src = synth.synth(160, 120, "mixture", 3)
    o = opts(score_tgt=80.0, max_pass=6)
    sc = OracleScorer(oracle)
    r, avif = H.search_image(src, o, scorer=sc)

    def probe(q):
        rgb = H.decode_rgb8(H.encode(src, q, o), 160, 120)
        return oracle.ssimu2_rgb8(src, rgb)

    want = z_search(probe, 80.0, 2.0, 6)
    assert (r.q, r.score, r.history()) == (want[0], want[1], want[2])
    assert r.num_pass == len(want[2]) == sc.n
    assert avif == H.encode(src, r.q, o) and r.size == len(avif)
    log = r.log.decode()
    assert f"Found q{r.q} (score {r.score:.2f}, {r.num_pass} passes)" in log      # measure.py parses "N passes"
    assert "Searching [tgt 80±2.0, speed 9, 8-bit]" in log and f"Compressed to {len(avif)} bytes" in log


def test_quality_bypass_and_alpha_input(oracle):
    rgba = synth.synth_rgba(96, 64, "noise", 2)
    r, avif = H.search_image(rgba, opts(quality=55), scorer=OracleScorer(oracle))
    assert r.num_pass == 0 and r.q == 55 and avif == H.encode(rgba, 55, opts())
    assert "Encoding [q55, speed 9, 8-bit]" in r.log.decode()
    # search on RGBA: alpha is encoded but never scored (io.zig:106-111, 654-663)
    sc = OracleScorer(oracle)
    r2, _ = H.search_image(rgba, opts(max_pass=3), scorer=sc)
    np.testing.assert_array_equal(sc.src, rgba[..., :3])
    assert r2.num_pass == sc.n >= 1


def test_batched_search_same_result_with_injected_oracle(oracle):
    src = synth.synth(128, 96, "edges", 5)
    o = opts(score_tgt=85.0, tolerance=1.0, max_pass=6)
    r1, a1 = H.search_image(src, o, scorer=OracleScorer(oracle))
    r4, a4 = H.search_image(src, o, batch_width=4, scorer=OracleScorer(oracle))
    assert (r4.q, r4.score, r4.history(), r4.num_pass) == (r1.q, r1.score, r1.history(), r1.num_pass)
    assert a4 == a1 and r4.device_passes <= r1.num_pass


def test_corpus_driver_without_a_gpu_records_errors_and_keeps_going(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    path = str(tmp_path / "c.csv")
    r = H.corpus_synth(3, 96, 64, opts=opts(max_pass=2), csv_path=path)
    assert r["ok"] == 0 and "Images: 0 ok, 0 no-output, 3 errors" in r["summary"]
    rows = list(csv.reader(open(path, newline="")))
    assert len(rows) == 4 and all(x[7] == "error" and "oavif_ssimu2_ctx_create" in x[8] for x in rows[1:])  # no CPU fallback


def _oracle_pair(oracle):
    def score_pair(src_rgb, y, u, v, depth, matrix, rgba):
        return oracle.ssimu2_rgb8(src_rgb, oracle.yuv444_to_rgb8(y, u, v, depth, matrix, rgba), oracle.BLUR_IIR, fast=True)
    return score_pair


def test_corpus_cpu_scored_arm_summary_and_trace(tmp_path, oracle):
    """The corpus driver with the injected CPU scorer (the arm bench tooling times beside the GPU): measure.py's CSV
    and summary text line for line (measure.py:178-269), the per-image trace, dynamic work sharing."""
    o = opts(max_pass=4, speed=10)
    path = str(tmp_path / "cpu.csv")
    r = H.corpus_synth(5, 160, 96, n_gpus=1, workers_per_gpu=3, opts=o, csv_path=path, score_pair=_oracle_pair(oracle))
    assert r["ok"] == 5 and r["errors"] == 0 and r["workers"] == 3, r
    rows = list(csv.reader(open(path, newline="")))
    assert rows[0] == ["Image", "Original Bytes", "Final Bytes", "Savings Bytes", "Savings %", "Encoding Time (ms)",
                       "Passes", "Status", "Error"]
    assert [x[0] for x in rows[1:]] == [f"synth_{i:05d}_k{i % 4}_160x96" for i in range(5)]     # image order, whoever ran it
    orig = sum(int(x[1]) for x in rows[1:])
    fin = sum(int(x[2]) for x in rows[1:])
    lines = r["summary"].splitlines()
    for want in ("Run Summary", "Images: 5 ok, 0 no-output, 0 errors", "Compression Totals",
                 f"Original total bytes: {orig} ({orig / 1024:.2f} KiB)", f"Final total bytes:    {fin} ({fin / 1024:.2f} KiB)",
                 f"Savings (bytes):      {orig - fin} ({(orig - fin) / 1024:.2f} KiB)",
                 f"% saved (overall):    {100.0 * (orig - fin) / orig:.2f}%", "Timing & Passes"):
        assert want in lines, (want, r["summary"])
    import statistics
    geo = (1.0 - statistics.geometric_mean([int(x[2]) / int(x[1]) for x in rows[1:]])) * 100.0
    assert f"% saved (geometric mean across files): {geo:.2f}%" in lines
    assert any(l.startswith("Input bytes throughput: ") and l.endswith("/s") for l in lines)
    passes = [int(x[6]) for x in rows[1:]]
    assert any(l.startswith(f"Average passes:        {sum(passes) / 5:.2f} ± ") and
               l.endswith(f"(max: {max(passes)}, min: {min(passes)})") for l in lines)
    assert sum(r["margin_hist"].values()) == 5 and r["scorer_device_ms"] == 0.0
    trace = list(csv.reader(open(path + ".trace.csv", newline="")))
    assert trace[0][:4] == ["Image", "Q", "Score", "Passes"] and len(trace) == 6
    for t, x in zip(trace[1:], rows[1:]):
        assert t[0] == x[0] and t[3] == x[6] and len(t[11].split()) == int(t[3]) and float(t[10]) >= 0.0
    # one worker, same per-image results: the shared counter changes who runs what, never what comes out
    r1 = H.corpus_synth(5, 160, 96, n_gpus=1, workers_per_gpu=1, opts=o, csv_path=path + ".1", score_pair=_oracle_pair(oracle))
    rows1 = list(csv.reader(open(path + ".1", newline="")))
    assert r1["ok"] == 5 and [(x[0], x[2], x[6]) for x in rows1] == [(x[0], x[2], x[6]) for x in rows]


def test_cli_reads_pam_and_ppm_and_reports_like_the_reference(tmp_path):
    import subprocess
    cli = os.path.join(os.path.dirname(H.LIB_PATH), "oavif-b200")
    img = synth.synth_rgba(40, 30, "noise", 1)
    pam = tmp_path / "a.pam"
    pam.write_bytes(b"P7\nWIDTH 40\nHEIGHT 30\nDEPTH 4\nMAXVAL 255\nTUPLTYPE RGB_ALPHA\nENDHDR\n" + img.tobytes())
    ppm = tmp_path / "a.ppm"
    ppm.write_bytes(b"P6\n# comment\n40 30\n255\n" + img[..., :3].tobytes())
    env = dict(os.environ, OAVIF_LIBAVIF=H.find_libavif())
    for f, kind in ((pam, "RGBA"), (ppm, "RGB")):
        out = tmp_path / (f.name + ".avif")
        p = subprocess.run([cli, "--tenbit", "0", "-q", "60", str(f), str(out)], capture_output=True, text=True, env=env)
        assert p.returncode == 0, p.stderr
        assert f"Read 40x30, {kind}, 8-bit, {f.stat().st_size} bytes" in p.stderr
        assert "Encoding [q60, speed 9, 8-bit]" in p.stderr and f"Compressed to {out.stat().st_size} bytes" in p.stderr
        px = img if kind == "RGBA" else np.ascontiguousarray(img[..., :3])
        assert out.read_bytes() == H.encode(px, 60, opts())
    bad = subprocess.run([cli, "--speed", "11", str(ppm), "x.avif"], capture_output=True, text=True, env=env)
    assert bad.returncode != 0 and "--speed must be between 0 and 10" in bad.stderr           # parse_args.zig:84
    miss = subprocess.run([cli, "--tolerance", "-3", str(ppm), "x.avif"], capture_output=True, text=True, env=env)
    assert miss.returncode != 0 and "Missing value" in miss.stderr                             # parse_args.zig:126


# ---- GPU ------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("case", [(256, 192, "mixture", 80.0, 2.0), (320, 200, "noise", 70.0, 1.0),
                                  (200, 300, "edges", 90.0, 2.0), (512, 384, "gradient", 85.0, 1.0)])
def test_gpu_scored_search_equals_oracle_scored_search(oracle, case):
    w, h, kind, tgt, tol = case
    src = synth.synth(w, h, kind, 11)
    o = opts(score_tgt=tgt, tolerance=tol, max_pass=6)
    want, bytes_cpu = H.search_image(src, o, scorer=OracleScorer(oracle))
    got, bytes_gpu = H.search_image(src, o, device=0)
    assert (got.q, got.num_pass, [q for q, _ in got.history()]) == (want.q, want.num_pass, [q for q, _ in want.history()])
    assert bytes_gpu == bytes_cpu
    for (_, a), (_, b) in zip(got.history(), want.history()):
        assert abs(a - b) <= 1e-4
    bat, bytes_bat = H.search_image(src, o, batch_width=4, device=0)
    assert (bat.q, bat.num_pass, bat.history()) == (got.q, got.num_pass, got.history()) and bytes_bat == bytes_gpu


@pytest.mark.gpu
def test_corpus_driver_csv_and_sharding(tmp_path):
    o = opts(max_pass=3, speed=10)
    csv1 = str(tmp_path / "one.csv")
    r1 = H.corpus_synth(6, 192, 128, n_gpus=1, workers_per_gpu=1, opts=o, csv_path=csv1)
    assert r1["ok"] == 6, r1
    rows = list(csv.reader(open(csv1, newline="")))
    assert rows[0] == ["Image", "Original Bytes", "Final Bytes", "Savings Bytes", "Savings %", "Encoding Time (ms)",
                       "Passes", "Status", "Error"]                                   # measure.py:180-192
    assert len(rows) == 7 and all(r[7] == "ok" for r in rows[1:])
    assert [r[0] for r in rows[1:]] == [f"synth_{i:05d}_k{i % 4}_192x128" for i in range(6)]
    assert "Throughput:" in r1["summary"] and "Average passes:" in r1["summary"]
    # more workers: same per-image results (only the timings differ)
    csv3 = str(tmp_path / "three.csv")
    r3 = H.corpus_synth(6, 192, 128, n_gpus=1, workers_per_gpu=3, opts=o, csv_path=csv3)
    rows3 = list(csv.reader(open(csv3, newline="")))
    assert r3["ok"] == 6 and [(r[0], r[2], r[6]) for r in rows3] == [(r[0], r[2], r[6]) for r in rows]
