"""CPU-side checks: the C-ABI library builds, loads and exports every symbol include/b200vmaf.h
declares (no compute calls without a GPU); model loading and SVR known answers; host logic."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from pqa2_b200 import _lib as L
from pqa2_b200 import engine, model as M, report

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    lib = L.load()
    hdr = open(os.path.join(ROOT, "include", "b200vmaf.h")).read()
    declared = set(re.findall(r"\b(bv_[a-z_0-9]+)\s*\(", hdr))
    declared -= {"bv_ctx", "bv_model"}
    assert declared, "no declarations parsed"
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in b200vmaf.h but not exported"
    assert set(L.EXPORTS) <= declared
    assert lib.bv_abi_version() == 2
    assert lib.bv_sizeof_frame_features() == C.sizeof(L.BvFrameFeatures)


def test_no_cpu_fallback_without_device():
    lib = L.load()
    if lib.bv_device_count() > 0:
        pytest.skip("a GPU is present")
    from pqa2_b200.extractor import BvError, FeatureExtractor
    with pytest.raises(BvError) as e:
        FeatureExtractor(320, 240)
    assert "no CUDA device" in str(e.value) or "CUDA" in str(e.value)


def test_bad_arguments_rejected_before_touching_cuda():
    lib = L.load()
    assert not lib.bv_create(0, 8, 8, 8, 420, L.FEAT_VMAF_INT, None)
    assert b"width/height" in lib.bv_last_error(None)
    assert not lib.bv_create(0, 640, 480, 9, 420, L.FEAT_VMAF_INT, None)
    assert not lib.bv_create(0, 640, 480, 8, 411, L.FEAT_VMAF_INT, None)
    assert not lib.bv_create(0, 640, 480, 8, 420, 0, None)


# ---- SVR pins computed from the reference's own model files ----
@pytest.mark.parametrize("name,feat,expect,clip", [
    ("vmaf_v0.6.1", [1, 0, 1, 1, 1, 1], 97.42804264, True),
    ("vmaf_v0.6.1", [0.9, 3.0, 0.5, 0.85, 0.92, 0.95], 72.57929533, True),
    ("vmaf_4k_v0.6.1", [1, 0, 1, 1, 1, 1], 100.8938, False),
    ("vmaf_4k_v0.6.1", [1, 0, 1, 1, 1, 1], 100.0, True),
    ("vmaf_4k_v0.6.1", [0.9, 3.0, 0.5, 0.85, 0.92, 0.95], 80.20082095, True),
    ("vmaf_b_v0.6.3", [1, 0, 1, 1, 1, 1], 97.98639, True),
    ("vmaf_float_v0.6.1", [1, 0, 1, 1, 1, 1], 97.42804264, True),
])
def test_svr_known_answers(name, feat, expect, clip):
    m = M.resolve_model(name)
    got = m.main.predict(np.array([feat], float), disable_clip=not clip)[0]
    assert abs(got - expect) < 5e-5 if expect != 100.0 else got == 100.0


def test_svr_host_matches_oracle_bit_exact():
    import oracle
    m = M.resolve_model("vmaf_v0.6.1").main
    rng = np.random.default_rng(0)
    feats = rng.uniform([0.3, 0, 0.1, 0.3, 0.4, 0.5], [1.0, 20, 1, 1, 1, 1], size=(64, 6))
    got = m.predict(feats, disable_clip=True)
    for i in range(64):
        assert got[i] == oracle.svr_predict(feats[i], m.slopes, m.intercepts, m.sv, m.coef, m.gamma, m.rho)


@pytest.mark.parametrize("name", ["vmaf_v0.6.1", "vmaf_float_v0.6.1", "vmaf_4k_v0.6.1", "vmaf_v0.6.1neg", "vmaf_b_v0.6.3"])
def test_svr_against_libsvm_itself(name):
    """A THIRD-PARTY pin of the fusion stage: scikit-learn vendors libsvm, so its NuSVR.predict runs libsvm's own
    svm_predict.  The fitted state of a dummy NuSVR is replaced by the parameters of the reference's model file (support
    vectors, coefficients, rho, gamma -- `models/*.json`, libsvm text under "model"), and libsvm's prediction on the
    normalised features, de-normalised as libvmaf's predict.c does, must equal the C oracle's and the C-ABI library's
    host evaluation (every support vector, every bootstrap model)."""
    svm = pytest.importorskip("sklearn.svm")
    import oracle
    vm = M.resolve_model(name)
    rng = np.random.default_rng(7)
    feats = rng.uniform([0.3, 0, 0.1, 0.3, 0.4, 0.5], [1.0, 20, 1, 1, 1, 1], size=(48, 6))
    for m in [vm.main] + list(vm.bootstrap)[:3]:
        n_sv, nf = m.sv.shape
        s = svm.NuSVR(kernel="rbf", gamma=m.gamma).fit(rng.random((max(50, n_sv), nf)), rng.random(max(50, n_sv)))
        s.support_vectors_ = np.ascontiguousarray(m.sv, np.float64)
        s.support_ = np.arange(n_sv, dtype=np.int32)
        s._n_support = np.array([n_sv, 0], dtype=np.int32)
        s.dual_coef_ = s._dual_coef_ = np.ascontiguousarray(m.coef.reshape(1, -1), np.float64)
        s.intercept_ = s._intercept_ = np.array([-m.rho])            # libsvm: sum - rho
        s._gamma = m.gamma
        s.shape_fit_ = (n_sv, nf)
        want = (s.predict(feats * m.slopes[1:] + m.intercepts[1:]) - m.intercepts[0]) / m.slopes[0]
        got_c = m.predict(feats, disable_clip=True)
        got_o = np.array([oracle.svr_predict(f, m.slopes, m.intercepts, m.sv, m.coef, m.gamma, m.rho) for f in feats])
        np.testing.assert_allclose(got_o, want, rtol=0, atol=1e-11)
        np.testing.assert_allclose(got_c, want, rtol=0, atol=1e-11)
        assert want.std() > 5                                        # un-clipped scores of random feature vectors: spread out


def test_model_files():
    names = M.available_models()
    for n in ["vmaf_v0.6.1", "vmaf_v0.6.1neg", "vmaf_4k_v0.6.1", "vmaf_4k_v0.6.1neg", "vmaf_b_v0.6.3",
              "vmaf_float_v0.6.1", "vmaf_float_v0.6.1neg", "vmaf_float_4k_v0.6.1", "vmaf_float_b_v0.6.3"]:
        assert n in names
    neg = M.resolve_model("vmaf_v0.6.1neg")
    assert neg.vif_enhn_gain_limit == 1.0 and neg.adm_enhn_gain_limit == 1.0
    b = M.resolve_model("vmaf_b_v0.6.3")
    assert len(b.bootstrap) == 20
    assert M.resolve_model("vmaf_float_v0.6.1").is_float and not M.resolve_model("vmaf_v0.6.1").is_float
    assert M.resolve_model(None).name == "vmaf_v0.6.1"
    assert M.resolve_model("vmaf_v0.6.1").main.metric_keys == [
        "integer_adm2", "integer_motion2", "integer_vif_scale0", "integer_vif_scale1", "integer_vif_scale2",
        "integer_vif_scale3"]
    with pytest.raises(FileNotFoundError):
        M.resolve_model("no_such_model")


def test_libsvm_sparse_rows():
    sv, coef, gamma, rho = M.parse_libsvm_text(
        "svm_type nu_svr\nkernel_type rbf\ngamma 0.04\nnr_class 2\ntotal_sv 2\nrho -1.5\nSV\n"
        "-4 1:0.5 2:0.25 4:-1.11e-16 6:2.22e-16 \n4 1:1 3:0.5 \n", 6)
    assert sv.shape == (2, 6) and sv[0, 2] == 0 and sv[0, 4] == 0 and sv[0, 3] == -1.11e-16 and sv[1, 2] == 0.5
    assert list(coef) == [-4, 4] and gamma == 0.04 and rho == -1.5


def test_motion2_rule_and_shards():
    assert engine.motion2_from_motion([0.0, 3.0, 1.0, 2.0]) == [0.0, 1.0, 1.0, 2.0]
    assert engine.shard_ranges(10, 3) == [(0, 3), (3, 6), (6, 10)]
    assert engine.shard_ranges(2, 8) == [(0, 1), (1, 2)]
    r = engine.shard_ranges(3600, 8)
    assert r[0] == (0, 450) and r[-1] == (3150, 3600) and all(a[1] == b[0] for a, b in zip(r, r[1:]))


def test_pooling_and_json_writer(tmp_path):
    frames = [{"frameNum": i, "metrics": {"vmaf": v, "integer_motion2": 0.5 * i}} for i, v in enumerate([90.0, 80.0, 70.0])]
    p = report.pooled_metrics(frames)
    assert p["vmaf"]["min"] == 70 and p["vmaf"]["max"] == 90 and p["vmaf"]["mean"] == 80
    assert abs(p["vmaf"]["harmonic_mean"] - (3 / (1 / 91 + 1 / 81 + 1 / 71) - 1)) < 1e-12
    out = tmp_path / "x_vmaf.json"
    report.write_libvmaf_json(str(out), frames, p, 123.456)
    import json
    d = json.loads(out.read_text())
    assert d["frames"][1]["frameNum"] == 1 and d["frames"][1]["metrics"]["vmaf"] == 80.0
    assert d["pooled_metrics"]["vmaf"]["mean"] == 80.0 and d["fps"] == 123.46
    assert '"vmaf": 80.000000' in out.read_text()
    csvp = tmp_path / "x.csv"
    report.write_frames_csv(str(csvp), frames)
    lines = csvp.read_text().splitlines()
    assert lines[0] == "Frame Number,integer_motion2,vmaf" and lines[2] == "1,0.5000,80.0000"


def test_result_csv_combined_csv_and_metadata(tmp_path):
    """Row f3: the per-test CSV, the combined CSV and the metadata file the reference's Results / History tabs
    produce and index (app/ui/tabs/results_tab.py:3518-3696, app/ui/tabs/analysis_tab.py:765-811)."""
    import csv
    import json
    from pqa2_b200 import report
    frames = [{"frameNum": i, "metrics": {"vmaf": 90.0 + i, "psnr_y": 40.125, "integer_motion2": 0.5}} for i in range(3)]
    data = {"frames": frames, "pooled_metrics": report.pooled_metrics(frames)}
    p = report.write_result_csv(str(tmp_path / "t.csv"), "T1", data, "/r/ref.y4m", "/r/dis.y4m", date="2025-01-01 00:00:00")
    rows = list(csv.reader(open(p)))
    assert rows[0] == ["Test Name", "Date", "VMAF Score", "PSNR Score", "SSIM Score"]
    assert rows[1] == ["T1", "2025-01-01 00:00:00", "91.0000", "40.1250", "N/A"]
    assert rows[2] == [] and rows[3] == ["Reference File", "/r/ref.y4m"] and rows[4] == ["Distorted File", "/r/dis.y4m"]
    assert rows[6] == ["Frame Number", "integer_motion2", "psnr_y", "vmaf"]
    assert rows[7] == ["0", "0.5000", "40.1250", "90.0000"] and len(rows) == 10
    p = report.write_combined_csv(str(tmp_path / "c.csv"), [
        {"test_name": "A", "timestamp": "2025-01-01 00:00:00", "vmaf_score": 91.0, "psnr_score": "a_psnr.txt",
         "ssim_score": None, "reference": "ref.y4m", "duration": "1.0s", "test_dir": "/x/A"}])
    rows = list(csv.reader(open(p)))
    assert rows[0][:5] == ["Test Name", "Date/Time", "VMAF Score", "PSNR Score", "SSIM Score"]
    assert rows[1] == ["A", "2025-01-01 00:00:00", "91.0000", "a_psnr.txt", "N/A", "ref.y4m", "1.0s", "/x/A"]
    res = {"vmaf_score": 91.0, "reference_video": "ref.y4m", "distorted_video": "dis.y4m", "psnr_score": "T_psnr.txt",
           "psnr_log": "/x/T_psnr.txt", "ssim_score": "Not Available", "ssim_log": None, "json_path": "/x/T_vmaf.json"}
    p = report.write_metadata_json(str(tmp_path / "T_metadata.json"), res,
                                   {"width": 1920, "height": 1080, "fps": 30.0, "frame_count": 3, "duration_seconds": 0.1},
                                   {"model": "vmaf_v0.6.1", "pool_method": "mean"}, "T")
    m = json.load(open(p))
    assert m["vmaf_score"] == 91.0 and m["json_result"] == "T_vmaf.json" and m["psnr_file"] == "T_psnr.txt"
    assert m["ssim_file"] is None and m["video_details"]["resolution"] == "1920x1080"
    assert m["analysis_settings"]["model"] == "vmaf_v0.6.1" and "os" in m["system_info"]


def test_adm_reciprocal_formula_is_exact():
    """The ADM kernels compute libvmaf's div_lookup entries (2^30 / d, truncated) as trunc(RN(1/d) * 2^30) in double
    instead of loading a 256 KB table; IEEE double arithmetic in numpy is the same arithmetic: all 32768 divisors."""
    import numpy as np
    d = np.arange(1, 32769, dtype=np.int64)
    assert np.array_equal(np.trunc((1.0 / d.astype(np.float64)) * 1073741824.0).astype(np.int64), 1073741824 // d)


def test_ffmpeg_stats_file_formats(tmp_path):
    """Row a9: the files the reference's `psnr` / `ssim` ffmpeg passes leave behind (app/vmaf_analyzer.py:1027-1034,
    :1057-1064) -- one line per frame, n from 1, FFmpeg's field names and precisions."""
    import math
    from pqa2_b200 import report
    areas = [1920 * 1080, 960 * 540, 960 * 540]
    p = tmp_path / "psnr.txt"
    out = report.write_ffmpeg_psnr_stats(str(p), [{"mse": [4.0, 1.0, 0.25], "areas": areas},
                                                    {"mse": [0.0, 0.0, 0.0], "areas": areas}], 8)
    l1, l2 = p.read_text().splitlines()
    avg = (4.0 * 4 + 1.0 + 0.25) / 6
    db = lambda m: 10 * math.log10(255 * 255 / m)                     # noqa: E731
    assert l1.split() == (f"n:1 mse_avg:{avg:.2f} mse_y:4.00 mse_u:1.00 mse_v:0.25 psnr_avg:{db(avg):.2f} "
                          f"psnr_y:{db(4.0):.2f} psnr_u:{db(1.0):.2f} psnr_v:{db(0.25):.2f}").split()
    assert l2.split()[:2] == ["n:2", "mse_avg:0.00"] and "psnr_avg:inf" in l2 and "psnr_y:inf" in l2
    assert out["mse_avg"] == pytest.approx(avg / 2)
    s = tmp_path / "ssim.txt"
    report.write_ffmpeg_ssim_stats(str(s), [{"ssim": [0.9, 0.96, 0.98], "weights": areas},
                                            {"ssim": [1.0, 1.0, 1.0], "weights": areas}])
    a1, a2 = s.read_text().splitlines()
    allv = (4 * 0.9 + 0.96 + 0.98) / 6                                # area-weighted = (4Y + U + V) / 6 at 4:2:0
    assert a1 == "n:1 Y:%f U:%f V:%f All:%f (%f)" % (0.9, 0.96, 0.98, allv, -10 * math.log10(1 - allv))
    assert a2 == "n:2 Y:1.000000 U:1.000000 V:1.000000 All:1.000000 (inf)"
