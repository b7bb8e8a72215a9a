"""CPU tests: the drop-in VMAFAnalyzer's conventions (no GPU needed for the error paths) and the
one-process-per-GPU host logic over gloo with world_size 2 (shards, lead-in, gather, motion2 across
the shard boundary)."""
import os
import socket

import numpy as np
import pytest

from pqa2_b200 import _lib as L
from pqa2_b200 import dist as D
from pqa2_b200 import engine, model as M, yuvio
from pqa2_b200.vmaf_analyzer import VMAFAnalyzer, VMAFAnalysisThread


def test_analyzer_surface_matches_reference():
    a = VMAFAnalyzer()
    for name in ("set_options_from_manager", "set_options_manager", "set_output_directory", "set_test_name",
                 "set_advanced_options", "terminate_analysis", "get_video_metadata", "analyze_videos"):
        assert callable(getattr(a, name))
    for sig in ("analysis_progress", "analysis_complete", "error_occurred", "status_update"):
        s = getattr(a, sig)
        assert hasattr(s, "connect") and hasattr(s, "emit")
    # reference defaults (app/vmaf_analyzer.py:25-39)
    assert (a.threads, a.pool_method, a.feature_subsample, a.psnr_enabled, a.ssim_enabled) == (4, "mean", 1, True, True)
    assert a.enable_motion_score is False and a.enable_temporal_features is False

    class Mgr:
        def get_setting(self, k):
            assert k == "vmaf"
            return {"threads": 8, "feature_subsample": 3, "pool_method": "harmonic_mean", "psnr_enabled": False}
    a.set_options_from_manager(Mgr())
    assert (a.threads, a.feature_subsample, a.pool_method, a.psnr_enabled, a.ssim_enabled) == (8, 3, "harmonic_mean", False, True)
    a.set_advanced_options(pool_method="min", feature_subsample=2)
    assert a.pool_method == "min" and a.feature_subsample == 2
    # signals are per instance
    b = VMAFAnalyzer()
    got = []
    a.error_occurred.connect(got.append)
    b.error_occurred.emit("x")
    assert got == []


def test_analyzer_never_raises_missing_file(tmp_path):
    a = VMAFAnalyzer()
    errs, done = [], []
    a.error_occurred.connect(errs.append)
    a.analysis_complete.connect(done.append)
    assert a.analyze_videos(str(tmp_path / "nope_ref.y4m"), str(tmp_path / "nope_dis.y4m")) is None
    assert errs and "Reference video not found" in errs[0] and not done


def test_analyzer_reports_engine_errors_instead_of_raising(tmp_path):
    """Without a GPU (this container) the engine error is emitted, not raised, and None comes back."""
    if L.load().bv_device_count() > 0:
        pytest.skip("a GPU is present")
    w, h = 64, 48
    fr = [[np.full((h, w), 100, np.uint8), np.full((h // 2, w // 2), 128, np.uint8), np.full((h // 2, w // 2), 128, np.uint8)]]
    r, d = tmp_path / "r.y4m", tmp_path / "d.y4m"
    yuvio.write_y4m(str(r), fr, w, h)
    yuvio.write_y4m(str(d), fr, w, h)
    a = VMAFAnalyzer()
    a.set_output_directory(str(tmp_path))
    a.set_test_name("T")
    errs = []
    a.error_occurred.connect(errs.append)
    assert a.analyze_videos(str(r), str(d)) is None
    assert errs and "CUDA" in errs[0]
    meta = a.get_video_metadata(str(r))
    assert meta["width"] == w and meta["height"] == h and meta["nb_frames"] == 1 and meta["pix_fmt"] == "yuv420p"
    assert a.get_video_metadata(str(tmp_path / "missing.y4m")) is None


def test_analysis_thread_forwards_signals(tmp_path):
    t = VMAFAnalysisThread(str(tmp_path / "a.y4m"), str(tmp_path / "b.y4m"))
    errs = []
    t.error_occurred.connect(errs.append)
    t.start()
    t.join(30)
    assert t.results is None and errs


# ---- one process per GPU, world_size 2 on gloo ---------------------------------------------------
class _FakeClip:
    width, height, bpc, chroma, fps = 64, 48, 8, 0, 30.0

    def __init__(self, n):
        self.nb_frames = n


def _fake_row(i, prev_seen):
    """Stand-in for the CUDA extractors: features are functions of the frame index; motion needs the
    previous frame (so a shard without its lead-in frame would get it wrong)."""
    motion = 0.0 if i == 0 else abs(np.sin(i * 1.7) - np.sin((i - 1) * 1.7)) * 5
    assert i == 0 or prev_seen, "shard started without its lead-in frame"
    v = [0.5 + 0.4 * np.cos(i * 0.3 + s) ** 2 for s in range(4)]
    return {"valid": L.FEAT_VMAF_INT, "raw": [0] * 64, "motion": float(motion), "vif": v,
            "adm2": 0.8 + 0.1 * np.sin(i) ** 2, "adm_scale": [0.9] * 4, "adm_num": [1.0] * 4, "adm_den": [1.0] * 4,
            "vif_num": v, "vif_den": [1.0] * 4, "psnr": [40.0, 0, 0], "ffssim": [0, 0, 0], "f_motion": 0.0,
            "f_vif": [0] * 4, "f_adm2": 0.0, "f_adm_scale": [0] * 4, "float_ssim": 0.0, "float_ms_ssim": 0.0}


def _fake_shard(src, model, opt, device, start, end, mask):
    lead = 1 if start > 0 else 0
    out, seen_prev = {}, False
    for i in range(start - lead, end):
        if i >= start:
            out[i] = _fake_row(i, seen_prev or i == 0)
        seen_prev = True
    return out


class _FakeSession:
    def __init__(self, rank):
        self.rank = rank

    def close(self):
        pass

    def analyze(self, clip, model, opt):
        if clip.nb_frames == 13:
            raise RuntimeError("bad clip")
        assert opt.devices == (self.rank,)
        return {"pooled_metrics": {"vmaf": {"mean": float(clip.nb_frames)}}, "n_frames": clip.nb_frames,
                "model": model.name, "rank": self.rank}


def _worker(rank, world, port, n, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        model = M.resolve_model("vmaf_v0.6.1")
        opt = engine.EngineOptions(svr_on_device=False)
        res = D.analyze_distributed(_FakeClip(n), model, opt, shard_fn=_fake_shard)
        t = D.max_over_ranks(10.0 + rank)
        # whole clips dealt over the ranks (configs[4]): summaries come back on rank 0 in input order, a failing clip
        # fills its own slot only
        clips = [_FakeClip(k) for k in (11, 12, 13, 14, 15)]
        batch = D.analyze_batch_distributed(clips, model, opt, device=rank, session=_FakeSession(rank), concurrency=1)
        # two clips in flight per rank, each worker with a session of its own: same summaries, same slots
        made = []
        engine.Engine = lambda: (made.append(1), _FakeSession(rank))[1]
        batch2 = D.analyze_batch_distributed(clips, model, opt, device=rank, concurrency=2)
        assert len(made) == 2 and (batch2 == batch if rank == 0 else batch2 is None)
        if rank == 0:
            assert [b.get("n_frames") for b in batch] == [11, 12, None, 14, 15] and "bad clip" in batch[2]["error"]
            assert [b["rank"] for b in batch if "rank" in b] == [0, 1, 1, 0]
            q.put(([fr["metrics"] for fr in res["frames"]], res["pooled_metrics"]["vmaf"], t))
        else:
            assert res is None and batch is None and t == 10.0 + world - 1
    finally:
        dist.destroy_process_group()


def _failing_shard(src, model, opt, device, start, end, mask):
    if start > 0:
        raise ValueError("decoder gave up")
    return _fake_shard(src, model, opt, device, start, end, mask)


def _worker_one_rank_fails(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        try:
            D.analyze_distributed(_FakeClip(9), M.resolve_model("vmaf_v0.6.1"), engine.EngineOptions(svr_on_device=False),
                                  shard_fn=_failing_shard)
            q.put((rank, "returned"))
        except Exception as e:            # noqa: BLE001
            q.put((rank, f"{type(e).__name__}: {e}"))
    finally:
        dist.destroy_process_group()


def test_a_failing_rank_does_not_hang_the_gather():
    """A shard that raises on one rank: every rank still reaches the gather, rank 0 reports which rank failed and why,
    the failing rank re-raises its own exception -- nobody waits for ever."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_one_rank_fails, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert got[1] == "ValueError: decoder gave up"
    assert got[0].startswith("RuntimeError: rank 1: ValueError: decoder gave up")


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("n", [7, 2])
def test_two_rank_gloo_matches_single_process(n):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    metrics, pooled, tmax = q.get(timeout=120)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert tmax == 11.0
    # single-process truth
    model = M.resolve_model("vmaf_v0.6.1")
    rows = [_fake_row(i, True) for i in range(n)]
    frames = engine.build_frames(rows, model, engine.EngineOptions(svr_on_device=False), None)
    assert len(metrics) == n
    for a, b in zip(metrics, frames):
        assert a == b["metrics"]                     # bit-identical, incl. motion2 across the shard boundary
    assert D.rank_range(7, 0, 2) == (0, 3) and D.rank_range(7, 1, 2) == (3, 7) and D.rank_range(1, 1, 2) == (1, 1)


def _write_mp4(cv2, path, frames_bgr, w, h, fourcc="mp4v"):
    wr = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*fourcc), 25, (w, h))
    if not wr.isOpened():
        pytest.skip(f"cv2 cannot encode {fourcc} here")
    for fr in frames_bgr:
        wr.write(fr)
    wr.release()


@pytest.mark.parametrize("decoder", ["av", "cv2"])
def test_container_decode(tmp_path, monkeypatch, decoder):
    """Row f2 (decode): compressed clips come in through the libavformat / libavcodec inside the cv2 wheel -- all three
    planes (`av`), or through cv2.VideoCapture when those libraries cannot be driven -- luma only (`cv2`)."""
    cv2 = pytest.importorskip("cv2")
    from pqa2_b200 import avdec, synth
    if decoder == "av" and not avdec.available():
        pytest.skip("the cv2 wheel's FFmpeg libraries are not usable here")
    if decoder == "cv2":
        monkeypatch.setattr(avdec, "available", lambda: False)
    w, h, n = 320, 176, 6
    path = str(tmp_path / "clip.mp4")
    lum = [synth.ref_luma(1, f, w, h) for f in range(n)]
    _write_mp4(cv2, path, [cv2.cvtColor(y, cv2.COLOR_GRAY2BGR) for y in lum], w, h)
    info = yuvio.probe(path)
    chroma = 420 if decoder == "av" else 400
    assert (info.width, info.height, info.bpc, info.chroma, info.nb_frames, info.decoder) == (w, h, 8, chroma, n, decoder)
    assert abs(info.fps - 25.0) < 1e-6
    r = yuvio.ClipReader(info)
    pl = r.alloc_planes(pinned=False)
    assert len(pl) == (3 if decoder == "av" else 1)
    for f in (0, 1, 4, 2):                       # sequential and seeking reads
        r.read_into(f, pl, decoder == "cv2")
        # limited-range luma of a lossy encode: 16 + 219/255 * gray, within a few code values on average
        want = 16.0 + lum[f].astype(np.float64) * (219.0 / 255.0)
        assert pl[0].shape == (h, w) and np.abs(pl[0] - want).mean() < 4.0
        if decoder == "av":                      # a grey picture: both chroma planes sit at 128
            assert pl[1].shape == (h // 2, w // 2) and abs(pl[1].mean() - 128) < 2 and abs(pl[2].mean() - 128) < 2
    # past the decodable end: the clip ends there (ffmpeg + libvmaf stop at the shorter input), no seek involved
    with pytest.raises(yuvio.EndOfClip):
        r.read_into(n + 3, pl, True)
    r.read_into(1, pl, True)                     # behind the cursor: reopened and decoded forward, never seeked
    assert np.abs(pl[0] - (16.0 + lum[1].astype(np.float64) * (219.0 / 255.0))).mean() < 4.0
    r.close()
    meta = VMAFAnalyzer().get_video_metadata(path)
    assert meta["width"] == w and meta["nb_frames"] == n
    assert meta["codec_name"] == "mpeg4"         # ffprobe's name of the stream's codec, not "rawvideo"


@pytest.mark.parametrize("fourcc,ext", [("mp4v", "mp4"), ("MJPG", "avi")])
def test_avdec_planes_are_the_decoders(tmp_path, fourcc, ext):
    """avdec against cv2's own use of the same libavcodec: the luma plane is identical to what VideoCapture hands back
    raw, and Y, U, V converted to BGR reproduce VideoCapture's BGR output to within the two conversions' rounding (a
    swapped or misaligned chroma plane would be off by tens of code values on this colourful clip).  MJPG decodes to a
    full-range `yuvj` format: same planes."""
    cv2 = pytest.importorskip("cv2")
    from pqa2_b200 import avdec
    if not avdec.available():
        pytest.skip("the cv2 wheel's FFmpeg libraries are not usable here")
    w, h, n = 208, 112, 5
    yy, xx = np.mgrid[0:h, 0:w]
    frames = []
    for f in range(n):
        b = (96 + 90 * np.sin(xx / 17.0 + f)).astype(np.uint8)
        g = (128 + 100 * np.cos(yy / 13.0 - f)).astype(np.uint8)
        r_ = ((xx * 255) // w).astype(np.uint8)
        frames.append(cv2.GaussianBlur(np.dstack([b, g, r_]), (5, 5), 1.5))
    path = str(tmp_path / f"clip.{ext}")
    _write_mp4(cv2, path, frames, w, h, fourcc)
    raw, bgr = cv2.VideoCapture(path), cv2.VideoCapture(path)
    raw.set(cv2.CAP_PROP_CONVERT_RGB, 0)
    with avdec.AvDecoder(path) as d:
        assert (d.width, d.height, d.bpc) == (w, h, 8) and d.chroma in (420, 422, 444)
        assert d.codec_name == {"mp4v": "mpeg4", "MJPG": "mjpeg"}[fourcc] and (d.fps_num, d.fps_den) == (25, 1)
        pl = [np.zeros(s, np.uint8) for s in d.plane_shapes()]
        k = 0
        while d.next(pl):
            ok, lum = raw.read()
            ok2, want = bgr.read()
            assert ok and ok2
            if lum.ndim == 2 or lum.shape[-1] == 1:
                np.testing.assert_array_equal(lum.reshape(-1, w)[:h], pl[0])
            u, v = (cv2.resize(c, (w, h), interpolation=cv2.INTER_LINEAR) for c in pl[1:])
            # chroma must carry the picture's colour: far from flat, and the reconstruction close to the decoded BGR
            assert pl[1].std() > 5 and pl[2].std() > 5
            ycc = np.dstack([pl[0], v, u])                  # cv2's YCrCb order
            if d.pix_fmt.startswith("yuvj"):
                got = cv2.cvtColor(ycc, cv2.COLOR_YCrCb2BGR).astype(np.int32)
            else:
                yl = np.clip((pl[0].astype(np.float32) - 16) * (255 / 219), 0, 255)
                cs = [np.clip((c.astype(np.float32) - 128) * (255 / 224) + 128, 0, 255) for c in (v, u)]
                got = cv2.cvtColor(np.dstack([yl] + cs).round().astype(np.uint8), cv2.COLOR_YCrCb2BGR).astype(np.int32)
            assert np.abs(got - want.astype(np.int32)).mean() < 3.0
            swapped = cv2.cvtColor(np.dstack([pl[0], u, v]), cv2.COLOR_YCrCb2BGR).astype(np.int32)
            assert np.abs(swapped - want.astype(np.int32)).mean() > 10.0
            k += 1
        assert k == n and not d.next(pl)


def test_libvmaf_filter_string_round_trip():
    """Rows a4 / a15: the reference's option builder (app/vmaf_analyzer.py:373-406) and the legacy call site's
    string (app/ui/tabs/results_tab.py:346-350) both map onto engine options."""
    from pqa2_b200 import options as O
    s = O.build_libvmaf_filter("/out/T_1_vmaf.json", "vmaf_v0.6.1", threads=8, feature_subsample=3)
    assert s == "libvmaf=log_path=/out/T_1_vmaf.json:log_fmt=json:model=version=vmaf_v0.6.1:n_threads=8:n_subsample=3"
    p = O.parse_libvmaf_filter(s)
    assert (p["model"], p["log_path"], p["log_fmt"], p["n_threads"], p["pool"]) == ("vmaf_v0.6.1", "/out/T_1_vmaf.json", "json", 8, "mean")
    assert p["options"].n_subsample == 3 and not p["options"].psnr and not p["options"].ssim
    s = O.build_libvmaf_filter("C:/r/x_vmaf.json", "C:/models/custom.json", pool_method="harmonic_mean",
                               enable_motion_score=True, enable_temporal_features=True)
    assert ":pool=harmonic_mean:psnr=1:ssim=1:" in s and s.count("feature=name=motion:enable=1") == 3
    p = O.parse_libvmaf_filter(s)
    assert p["model"] == "C:/models/custom.json" and p["log_path"] == "C:/r/x_vmaf.json" and p["pool"] == "harmonic_mean"
    assert p["options"].psnr and p["options"].ssim and "vif_scale0" in p["features"] and "adm2" in p["features"]
    legacy = O.parse_libvmaf_filter("libvmaf=log_fmt=json:log_path=/tmp/vmaf.json:psnr=1:ssim=1:model_path=/m/vmaf_v0.6.1.json")
    assert legacy["model"] == "/m/vmaf_v0.6.1.json" and legacy["options"].psnr and legacy["options"].ssim
    assert M.resolve_model(O.parse_libvmaf_filter("libvmaf=model=version=vmaf_4k_v0.6.1")["model"]).name == "vmaf_4k_v0.6.1"


def test_column_pooling_equals_the_per_frame_walk():
    """engine._pool_column (numpy, sequential cumsum) == report.pool (libvmaf's left-to-right double sums), bit for bit."""
    import numpy as np
    from pqa2_b200 import engine, report
    rng = np.random.default_rng(3)
    for n in (1, 2, 17, 513):
        col = rng.uniform(0.0, 100.0, n)
        assert engine._pool_column(col) == report.pool(col.tolist())
    assert engine._pool_column(np.zeros(0)) == report.pool([])


def test_weighted_shard_ranges():
    """engine.shard_ranges with per-shard rates: contiguous, complete, proportional; degenerate weights fall back to equal."""
    r = engine.shard_ranges(3600, 8, [24, 24, 24, 24, 36, 36, 36, 36])
    assert r[0][0] == 0 and r[-1][1] == 3600 and all(a[1] == b[0] for a, b in zip(r, r[1:]))
    lens = [b - a for a, b in r]
    assert lens[:4] == [360] * 4 and lens[4:] == [540] * 4
    assert engine.shard_ranges(10, 3, [1, 0, 1]) == engine.shard_ranges(10, 3)
    assert engine.shard_ranges(5, 2, [1.0, 1.0]) == [(0, 2), (2, 5)] or engine.shard_ranges(5, 2, [1.0, 1.0]) == [(0, 3), (3, 5)]
    assert D.rank_range(3600, 5, 8, weights=[24, 24, 24, 24, 36, 36, 36, 36]) == (1980, 2520)


@pytest.mark.parametrize("bpc,w,h", [(10, 208, 120), (8, 161, 97), (12, 176, 144)])
def test_avdec_is_bit_exact_on_raw_streams(tmp_path, bpc, w, h):
    """A .y4m file through libavformat's yuv4mpegpipe demuxer + rawvideo decoder must hand back exactly the samples that
    were written: pins the plane copy (line sizes, 16-bit little-endian samples, odd chroma sizes) and shows that what
    ffmpeg would deliver for such a file equals what this repo's own raw reader delivers."""
    from pqa2_b200 import avdec, synth
    if not avdec.available():
        pytest.skip("the cv2 wheel's FFmpeg libraries are not usable here")
    n = 3
    frames = [synth.frame_pair(6, f, w, h, bpc)[1] for f in range(n)]
    path = str(tmp_path / "clip.y4m")
    yuvio.write_y4m(path, frames, w, h, bpc, (24, 1))
    info = yuvio.probe(path)
    rd = yuvio.ClipReader(info)
    own = rd.alloc_planes(pinned=False)
    with avdec.AvDecoder(path) as d:
        assert (d.width, d.height, d.bpc, d.chroma, d.codec_name) == (w, h, bpc, 420, "rawvideo") and (d.fps_num, d.fps_den) == (24, 1)
        pl = [np.zeros(s, np.uint8 if bpc == 8 else np.uint16) for s in d.plane_shapes()]
        for f in range(n):
            assert d.next(pl)
            rd.read_into(f, own)
            for k in range(3):
                np.testing.assert_array_equal(pl[k], frames[f][k])
                np.testing.assert_array_equal(pl[k], own[k])
        assert not d.next(pl)
    rd.close()
