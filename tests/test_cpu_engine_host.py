"""Host side of the engine without a GPU: `engine.analyze` drives a stand-in for the CUDA context (same methods as
`extractor.FeatureExtractor`) and the test checks what the host loop is responsible for -- frame flags (first frame,
lead-in, n_subsample), launch-group kicks, shard ranges with lead-in frames, the motion2 rule over shard borders, the SVR
(host libsvm path of the C library), pooling -- against a single-shard run and against direct arithmetic.
Reference behaviour restated: libvmaf frame flow as reached from app/vmaf_analyzer.py:417."""
import ctypes as C

import numpy as np
import pytest

from pqa2_b200 import _lib as L
from pqa2_b200 import engine, model as M, report


class FakeExtractor:
    """Features are functions of the frame's content (here: of its first pixel = the global frame index)."""
    instances = []

    def __init__(self, width, height, bpc, chroma, features, device=0, **kw):
        self.batch = kw.get("batch_frames") or 32
        self.chroma, self.features = chroma, features
        self.frames, self.kicks, self.kw = [], [], kw
        FakeExtractor.instances.append(self)

    batch_frames = property(lambda self: self.batch)
    kernel_launches = 0

    def submit(self, frame_index, ref_planes, dis_planes, flags=0):
        self.frames.append((int(frame_index), int(ref_planes[0][0, 0]), int(flags)))

    def kick(self):
        self.kicks.append(len(self.frames))

    def wait_uploads(self):
        pass

    def flush(self):
        pass

    def reset(self):
        self.frames.clear()

    def cancel(self):
        pass

    def close(self):
        pass

    in_flight = 40           # frames still "on the GPU" when the last one has been submitted (the drain the host overlaps)

    def frames_done(self):
        return max(0, len(self.frames) - self.in_flight)

    def fetch(self, first=0, count=None):
        count = len(self.frames) - first if count is None else count
        full = self._all()
        arr = (L.BvFrameFeatures * count)()
        for k in range(count):
            C.memmove(C.byref(arr[k]), C.byref(full[first + k]), C.sizeof(L.BvFrameFeatures))
        return arr

    def _all(self):
        arr = (L.BvFrameFeatures * len(self.frames))()
        prev = None
        for k, (idx, content, flags) in enumerate(self.frames):
            f = arr[k]
            f.frame_index, f.flags = idx, flags
            lead = bool(flags & L.FRAME_LEAD_IN)
            spatial = not (flags & (L.FRAME_LEAD_IN | L.FRAME_SKIP_SPATIAL))
            valid = 0
            if not lead:
                f.motion = 0.0 if (flags & L.FRAME_FIRST) or prev is None else abs(np.sin(content * 1.7) - np.sin(prev * 1.7)) * 5
                valid |= L.FEAT_MOTION
            if spatial:
                for s in range(4):
                    f.vif_scale[s] = 0.5 + 0.4 * np.cos(content * 0.3 + s) ** 2
                    f.adm_scale[s] = 0.9
                f.adm2 = 0.8 + 0.1 * np.sin(content) ** 2
                f.psnr_y = 30.0 + content % 7
                valid |= L.FEAT_VIF | L.FEAT_ADM | L.FEAT_PSNR_Y
                if self.chroma not in (0, 400) and self.features & L.FEAT_PSNR_UV:
                    f.psnr_cb, f.psnr_cr = 40.0 + content % 5, 41.0 + content % 3
                    valid |= L.FEAT_PSNR_UV
            if self.features & L.FEAT_FLOAT_MOTION and not lead:
                f.f_motion = 0.0 if (flags & L.FRAME_FIRST) or prev is None else 1.0 + (content % 4)
                valid |= L.FEAT_FLOAT_MOTION
            f.valid_mask = valid
            prev = content
        return arr


class Clip(engine.FrameSource):
    width, height, bpc, chroma, fps = 64, 48, 8, 0, 30.0

    def __init__(self, n):
        self.nb_frames = n

    def read_into(self, i, ref_planes, dis_planes, luma_only):
        ref_planes[0][...] = i % 251
        dis_planes[0][...] = (i * 3) % 251


@pytest.fixture()
def fake(monkeypatch):
    FakeExtractor.instances = []
    monkeypatch.setattr(engine, "FeatureExtractor", FakeExtractor)
    monkeypatch.setattr(engine, "pinned_empty", lambda shape, dtype: np.empty(shape, dtype))
    return FakeExtractor


def _opt(**kw):
    return engine.EngineOptions(svr_on_device=False, **kw)


def test_flags_kicks_and_shards(fake):
    model = M.resolve_model("vmaf_v0.6.1")
    n = 75
    one = engine.analyze(Clip(n), model, _opt(psnr=True, devices=(0,)))
    (fx,) = fake.instances
    assert [i for i, _, _ in fx.frames] == list(range(n))
    assert fx.frames[0][2] & L.FRAME_FIRST and not any(fl & L.FRAME_FIRST for _, _, fl in fx.frames[1:])
    assert not any(fl & (L.FRAME_LEAD_IN | L.FRAME_SKIP_SPATIAL) for _, _, fl in fx.frames)
    assert fx.kicks == [8, n]                                # a short first group; no tail split at <= 1440p (32-frame groups);
                                                             # the last partial group is started before the early hand-over

    fake.instances.clear()
    three = engine.analyze(Clip(n), model, _opt(psnr=True, devices=(0, 0, 0)))
    assert len(fake.instances) == 3
    starts = sorted(fx.frames[0][0] for fx in fake.instances)
    assert starts == [0, 24, 49]                              # shard 0 from frame 0; the others begin with their lead-in frame
    for fx in fake.instances:
        first_idx, _, fl0 = fx.frames[0]
        assert fl0 & L.FRAME_FIRST
        assert bool(fl0 & L.FRAME_LEAD_IN) == (first_idx != 0)
        assert not any(fl & L.FRAME_LEAD_IN for _, _, fl in fx.frames[1:])
    assert [fr["metrics"] for fr in one["frames"]] == [fr["metrics"] for fr in three["frames"]]
    assert one["pooled_metrics"] == three["pooled_metrics"] == report.pooled_metrics(one["frames"])

    # motion2[i] = min(motion[i], motion[i+1]), last frame keeps its own; motion[0] = 0
    m = [fr["metrics"]["integer_motion"] for fr in one["frames"]]
    m2 = [fr["metrics"]["integer_motion2"] for fr in one["frames"]]
    assert m[0] == 0.0 and m2 == [min(m[i], m[i + 1]) if i + 1 < n else m[i] for i in range(n)]
    # the score is the C library's host SVR on (adm2, motion2, vif0..3) in the model's order
    fr = one["frames"][10]["metrics"]
    x = np.array([[fr["integer_adm2"], fr["integer_motion2"]] + [fr[f"integer_vif_scale{s}"] for s in range(4)]])
    assert model.main.predict(x)[0] == fr["vmaf"]


def test_n_subsample_scores_every_nth_frame_but_motion_sees_all(fake):
    model = M.resolve_model("vmaf_v0.6.1")
    res = engine.analyze(Clip(20), model, _opt(n_subsample=4))
    (fx,) = fake.instances
    assert [bool(fl & L.FRAME_SKIP_SPATIAL) for _, _, fl in fx.frames] == [i % 4 != 0 for i in range(20)]
    assert [fr["frameNum"] for fr in res["frames"]] == [0, 4, 8, 12, 16]
    assert all("integer_motion2" in fr["metrics"] and "vmaf" in fr["metrics"] for fr in res["frames"])


def test_tail_split_only_for_reduced_group_sizes(fake, monkeypatch):
    model = M.resolve_model("vmaf_v0.6.1")
    engine.analyze(Clip(100), model, _opt(batch_frames=16))
    (fx,) = fake.instances
    # first kick after 8 frames, then when 8 and 4 frames are left (2160p-style 16-frame groups), then the last partial
    # group before the rows that are already complete are handed over
    assert fx.kicks == [8, 92, 96, 100]


def test_long_clips_are_built_in_steps_while_frames_are_still_submitted(fake, monkeypatch):
    """Every 512 submitted frames the rows that are complete go to the frame builder (the submitting thread waits for ring
    slots most of the time anyway); the seams between the blocks -- motion2 needs the next frame -- and the pooled report
    must equal the one-pass build of the same rows."""
    model = M.resolve_model("vmaf_v0.6.1")
    calls = []
    real = engine._build_frames_block
    monkeypatch.setattr(engine, "_build_frames_block", lambda *a, **k: (calls.append(len(a[0])), real(*a, **k))[1])
    n = 1700
    one = engine.analyze(Clip(n), model, _opt(psnr=True, n_subsample=2, devices=(0,), contexts_per_device=1))
    # a hand-over once 512 more frames than at the last one are submitted (at a group boundary): 512, 992 and 1472 submitted
    # -> 472, 952, 1432 complete (40 "in flight"); then the one before the drain (1660 complete) and the rest.  Each block
    # starts one frame early: the last frame of a block waits for its successor's motion
    assert calls == [472, 952 - 471, 1432 - 951, 1660 - 1431, 1700 - 1659]
    fake.instances.clear()
    calls.clear()
    two = engine.analyze(Clip(n), model, _opt(psnr=True, n_subsample=2, devices=(0, 0), contexts_per_device=1,
                                              dynamic_chunk=0))
    assert calls == [n]
    assert one["frames"] == two["frames"] and one["pooled_metrics"] == two["pooled_metrics"]
    assert [fr["frameNum"] for fr in one["frames"]] == list(range(0, n, 2))
    # the default for a clip of this length: three contexts side by side on the GPU, chunks of n // 12 frames from a shared
    # counter, the waiting thread building the finished prefix -- same log
    fake.instances.clear()
    calls.clear()
    auto = engine.analyze(Clip(n), model, _opt(psnr=True, n_subsample=2, devices=(0,)))
    assert len(fake.instances) == 3 and sum(calls) >= n and len(calls) >= 1
    starts = sorted({fx.frames[0][0] for fx in fake.instances})
    assert auto["frames"] == one["frames"] and auto["pooled_metrics"] == one["pooled_metrics"]
    short = engine.analyze(Clip(1000), model, _opt(devices=(0,)))
    assert len(fake.instances) == 4                                       # below 1536 frames: one context
    assert len(short["frames"]) == 1000


@pytest.mark.parametrize("n,sub,rng", [(200, 1, None), (131, 3, None), (300, 1, (17, 260)), (105, 1, None), (90, 1, None)])
def test_frames_built_during_the_drain_equal_the_one_pass_build(fake, monkeypatch, n, sub, rng):
    """A single-shard analysis hands the rows that are complete when the last frame has been submitted to the frame
    builder while the GPU drains (engine.analyze `build_early`), and builds the rest afterwards: frames, motion2 across
    the seam, the optional columns and the pooled report must equal the one-pass build (three shards never take the
    early path), also with n_subsample and a frame range; too few complete rows -> the one-pass build."""
    model = M.resolve_model("vmaf_v0.6.1")
    calls = []
    real = engine._build_frames_block
    monkeypatch.setattr(engine, "_build_frames_block", lambda *a, **k: (calls.append(len(a[0])), real(*a, **k))[1])
    one = engine.analyze(Clip(n), model, _opt(psnr=True, n_subsample=sub, devices=(0,)), frame_range=rng)
    span = n if rng is None else rng[1] - rng[0]
    early = span - FakeExtractor.in_flight >= engine._EARLY_MIN
    assert len(calls) == (2 if early else 1)
    if early:
        assert calls[0] == span - FakeExtractor.in_flight and calls[1] == FakeExtractor.in_flight + 1
    fake.instances.clear()
    calls.clear()
    three = engine.analyze(Clip(n), model, _opt(psnr=True, n_subsample=sub, devices=(0, 0, 0)), frame_range=rng)
    assert len(calls) == 1
    assert one["frames"] == three["frames"] and one["pooled_metrics"] == three["pooled_metrics"]
    assert [fr["frameNum"] for fr in one["frames"]] == [i for i in range(*(rng or (0, n))) if i % sub == 0]
    assert list(one["pooled_metrics"]) == list(three["pooled_metrics"]) and "psnr_y" in one["pooled_metrics"]


def test_frame_range_and_empty_clip(fake):
    model = M.resolve_model("vmaf_v0.6.1")
    res = engine.analyze(Clip(40), model, _opt(), frame_range=(10, 25))
    assert [fr["frameNum"] for fr in res["frames"]] == list(range(10, 25))
    (fx,) = fake.instances
    # a frame range is scored like the trimmed clip: its first frame has no predecessor (libvmaf index 0: motion = 0)
    assert fx.frames[0][0] == 10 and fx.frames[0][2] & L.FRAME_FIRST and not fx.frames[0][2] & L.FRAME_LEAD_IN
    assert res["frames"][0]["metrics"]["integer_motion"] == 0.0
    fake.instances.clear()
    two = engine.analyze(Clip(40), model, _opt(devices=(0, 0)), frame_range=(10, 25))
    assert [fr["metrics"] for fr in two["frames"]] == [fr["metrics"] for fr in res["frames"]]
    assert sorted(fx.frames[0][0] for fx in fake.instances) == [10, 16]          # only the second shard gets a lead-in (frame 16)
    with pytest.raises(ValueError):
        engine.analyze(Clip(0), model, _opt())


class ChromaClip(Clip):
    chroma = 420

    def read_into(self, i, ref_planes, dis_planes, luma_only):
        for p in ref_planes:
            p[...] = i % 251
        for p in dis_planes:
            p[...] = (i * 3) % 251


def test_psnr_columns_and_feature_motion_adds_the_float_extractor(fake):
    """`psnr=1` (app/vmaf_analyzer.py:385) reaches libvmaf through FFmpeg's filter as the psnr extractor with
    enable_chroma=false: psnr_y only; the extractor's own default (psnr_chroma) adds psnr_cb /
    psnr_cr for pictures with chroma (the CSV export of results_tab.py:3009-3026 takes its header from these keys).
    `feature=name=motion` (:388-402) adds the float motion extractor's `motion` / `motion2` next to the integer model's."""
    model = M.resolve_model("vmaf_v0.6.1")
    plain = engine.analyze(ChromaClip(4), model, _opt(psnr=True))
    assert "psnr_y" in plain["frames"][0]["metrics"] and "psnr_cb" not in plain["frames"][0]["metrics"]
    fake.instances.clear()
    res = engine.analyze(ChromaClip(12), model, _opt(psnr_chroma=True, float_motion=True))
    m = res["frames"][5]["metrics"]
    assert {"psnr_y", "psnr_cb", "psnr_cr", "motion", "motion2", "integer_motion", "integer_motion2", "vmaf"} <= set(m)
    fm = [fr["metrics"]["motion"] for fr in res["frames"]]
    assert fm[0] == 0.0 and [fr["metrics"]["motion2"] for fr in res["frames"]] == \
        [min(fm[i], fm[i + 1]) if i + 1 < 12 else fm[i] for i in range(12)]
    assert set(res["pooled_metrics"]) >= {"psnr_cb", "psnr_cr", "motion2"}
    # luma-only pictures: no chroma PSNR, as libvmaf for gray input
    fake.instances.clear()
    res = engine.analyze(Clip(4), model, _opt(psnr_chroma=True))
    assert "psnr_y" in res["frames"][0]["metrics"] and "psnr_cb" not in res["frames"][0]["metrics"]


def test_failed_shard_stops_its_siblings_and_the_error_surfaces(fake, monkeypatch):
    class Boom(Clip):
        def read_into(self, i, ref_planes, dis_planes, luma_only):
            if i == 30:
                raise OSError("disk gone")
            super().read_into(i, ref_planes, dis_planes, luma_only)

    with pytest.raises(OSError):
        engine.analyze(Boom(4000), M.resolve_model("vmaf_v0.6.1"), _opt(devices=(0, 0)))
    assert max(len(fx.frames) for fx in fake.instances) < 2000           # the healthy shard did not run to its end


def test_file_source_reader_threads_fill_the_ring_in_order(fake, tmp_path):
    """Raw clips that cannot be mapped + registered (no GPU here; small files) go through the pinned ring filled by
    several reader threads: every frame must reach the extractor exactly once, in order, with its own content."""
    from pqa2_b200 import yuvio
    w, h, n = 64, 48, 150
    def frames(mul):
        for i in range(n):
            yield [np.full((h, w), (i * mul) % 251, np.uint8), np.full((h // 2, w // 2), 7, np.uint8),
                   np.full((h // 2, w // 2), 9, np.uint8)]
    rp, dp = str(tmp_path / "r.y4m"), str(tmp_path / "d.y4m")
    yuvio.write_y4m(rp, frames(1), w, h)
    yuvio.write_y4m(dp, frames(3), w, h)
    src = engine.FileSource(yuvio.probe(rp), yuvio.probe(dp))
    assert not src.zero_copy and src.parallel_reads
    res = engine.analyze(src, M.resolve_model("vmaf_v0.6.1"), _opt(batch_frames=8, reader_threads=4, devices=(0, 0)))
    src.release()
    got = sorted((idx, content) for fx in fake.instances for idx, content, fl in fx.frames if not fl & L.FRAME_LEAD_IN)
    assert got == [(i, i % 251) for i in range(n)]
    assert [fr["frameNum"] for fr in res["frames"]] == list(range(n))


def test_session_keeps_contexts_per_geometry_and_retain_drops_the_others(fake):
    model = M.resolve_model("vmaf_v0.6.1")
    closed = []
    fake.close = lambda self: closed.append(self)
    with engine.Engine() as sess:
        sess.analyze(Clip(10), model, _opt())
        sess.analyze(Clip(12), model, _opt())                    # same geometry: the context is reused, not rebuilt
        assert len(fake.instances) == 1

        class Small(Clip):
            width, height = 48, 32
        sess.analyze(Small(6), model, _opt())
        assert len(fake.instances) == 2 and not closed
        sess.retain(48, 32, 8)                                   # what VMAFAnalyzer does before the next analysis
        assert closed == [fake.instances[0]] and list(k[2:5] for k in sess._fx) == [(48, 32, 8)]


def test_long_clip_on_several_devices_is_dealt_in_chunks(fake):
    """>= 4 chunks per device: the devices pull chunks from a shared counter; every chunk after the first starts with its
    lead-in frame; the log is identical to the single-device one (motion2 across every chunk border)."""
    model = M.resolve_model("vmaf_v0.6.1")
    n = 400
    one = engine.analyze(Clip(n), model, _opt(devices=(0,)))
    fake.instances.clear()
    dyn = engine.analyze(Clip(n), model, _opt(devices=(0, 0), dynamic_chunk=40))
    assert [fr["metrics"] for fr in dyn["frames"]] == [fr["metrics"] for fr in one["frames"]]
    assert len(fake.instances) == 2                                  # one context per device, reused from chunk to chunk
    fixed = engine.analyze(Clip(n), model, _opt(devices=(0, 0), dynamic_chunk=0))
    assert [fr["metrics"] for fr in fixed["frames"]] == [fr["metrics"] for fr in one["frames"]]


def test_container_pair_is_decoded_in_step(fake, tmp_path):
    """Compressed inputs: one sequential decoder per file (each with libavcodec's frame threads: decoding the two files on
    two host threads was measured, +5 %, and left out); every frame pair must reach the extractor in order and in step,
    and a distorted stream that is shorter than its container claims ends the analysis there (ffmpeg + libvmaf stop at
    the shorter input)."""
    cv2 = pytest.importorskip("cv2")
    from pqa2_b200 import yuvio
    w, h = 64, 48

    def write(path, levels):
        wr = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"mp4v"), 25, (w, h))
        if not wr.isOpened():
            pytest.skip("cv2 cannot encode mp4v here")
        for v in levels:
            wr.write(np.full((h, w, 3), v, np.uint8))
        wr.release()

    n = 20
    ref_p, dis_p = str(tmp_path / "ref.mp4"), str(tmp_path / "dis.mp4")
    write(ref_p, [20 + 10 * i for i in range(n)])
    write(dis_p, [25 + 10 * i for i in range(n)])
    ri, di = yuvio.probe(ref_p), yuvio.probe(dis_p)
    assert ri.decoder != "raw" and ri.nb_frames == n
    src = engine.FileSource(ri, di)
    assert src.sequential and not src.parallel_reads
    seen = []
    real_submit = FakeExtractor.submit

    def submit(self, frame_index, ref_planes, dis_planes, flags=0):
        seen.append((int(frame_index), int(ref_planes[0][h // 2, w // 2]), int(dis_planes[0][h // 2, w // 2])))
        real_submit(self, frame_index, ref_planes, dis_planes, flags)

    FakeExtractor.submit = submit
    try:
        res = engine.analyze(src, M.resolve_model("vmaf_v0.6.1"), _opt(devices=(0,)))
        assert len(fake.instances) == 1 and len(res["frames"]) == n
        assert [s[0] for s in seen] == list(range(n))
        for idx, r, d in seen:
            # limited-range luma of a flat grey picture: 16 + 219 / 255 * level, within a couple of code values
            assert abs(r - (16 + (20 + 10 * idx) * 219 / 255)) < 3 and abs(d - (16 + (25 + 10 * idx) * 219 / 255)) < 3
        # the distorted container claims more frames than it holds
        di.nb_frames = ri.nb_frames = n + 7
        seen.clear()
        cut = engine.analyze(engine.FileSource(ri, di), M.resolve_model("vmaf_v0.6.1"), _opt())
        assert len(cut["frames"]) == n and [s[0] for s in seen] == list(range(n))
    finally:
        FakeExtractor.submit = real_submit
