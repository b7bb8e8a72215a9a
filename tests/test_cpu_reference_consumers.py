"""The reference's OWN consumer code on this engine's output.

`/root/reference/app/vmaf_analyzer.py::_parse_vmaf_results` (:628-980) is what turns libvmaf's JSON log and the two
FFmpeg stats files into the results dict the GUI shows; `results_tab.py:3009-3026` derives the CSV header from the
first frame's metric keys.  Here both are run, unmodified, on files written by pqa2_b200.report, and their results are
compared with what pqa2_b200.vmaf_analyzer returns for the same log.  PyQt5 is stubbed (absent in this image);
`/root/reference` does not travel to the GPU box, so the module skips there."""
import importlib
import json
import os
import sys
import types

import numpy as np
import pytest

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isfile(os.path.join(REF, "app", "vmaf_analyzer.py")),
                                reason="/root/reference is not present (GPU box): the reference's parser cannot be imported")


@pytest.fixture()
def ref_module(monkeypatch):
    """app.vmaf_analyzer of the reference, imported with a stand-in PyQt5.QtCore."""
    class Signal:
        def __init__(self, *a):
            self.sent = []

        def __set_name__(self, owner, name):
            self.name = "_sig_" + name

        def __get__(self, obj, owner=None):
            if obj is None:
                return self
            return obj.__dict__.setdefault(self.name, Signal())

        def emit(self, *a):
            self.sent.append(a)

        def connect(self, fn):
            pass

    qtcore = types.ModuleType("PyQt5.QtCore")
    qtcore.QObject = type("QObject", (), {"__init__": lambda self, *a, **k: None})
    qtcore.pyqtSignal = Signal
    pyqt = types.ModuleType("PyQt5")
    pyqt.QtCore = qtcore
    monkeypatch.setitem(sys.modules, "PyQt5", pyqt)
    monkeypatch.setitem(sys.modules, "PyQt5.QtCore", qtcore)
    monkeypatch.syspath_prepend(REF)
    for k in [k for k in sys.modules if k == "app" or k.startswith("app.")]:
        monkeypatch.delitem(sys.modules, k)
    mod = importlib.import_module("app.vmaf_analyzer")
    yield mod
    for k in [k for k in sys.modules if k == "app" or k.startswith("app.")]:
        sys.modules.pop(k, None)


def _engine_log(n=5, chroma=False):
    """A log as engine.analyze builds it, from hand-made feature rows (no GPU needed): integer model + psnr=1 + ssim=1."""
    from pqa2_b200 import _lib as L
    from pqa2_b200 import engine, model as M
    rows = engine.Rows(n)
    a = rows.arr
    rng = np.random.default_rng(5)
    a["valid_mask"] = L.FEAT_VMAF_INT | L.FEAT_PSNR_Y | (L.FEAT_PSNR_UV if chroma else 0) | L.FEAT_FLOAT_SSIM
    a["motion"] = rng.uniform(0, 6, n)
    a["adm2"] = rng.uniform(0.8, 1.0, n)
    a["adm_scale"] = rng.uniform(0.8, 1.0, (n, 4))
    a["vif_scale"] = rng.uniform(0.4, 1.0, (n, 4))
    a["psnr_y"], a["psnr_cb"], a["psnr_cr"] = rng.uniform(30, 45, n), rng.uniform(35, 48, n), rng.uniform(35, 48, n)
    a["float_ssim"] = rng.uniform(0.9, 1.0, n)
    rows.present[:] = True
    model = M.resolve_model("vmaf_v0.6.1")
    opt = engine.EngineOptions(psnr=True, psnr_chroma=chroma, ssim=True, svr_on_device=False)
    pooled_out = {}
    frames = engine.build_frames(rows, model, opt, None, pooled_out)
    return frames, pooled_out["pooled"]


def test_reference_parser_reads_our_log_and_stats_files(ref_module, tmp_path):
    from pqa2_b200 import report
    frames, pooled = _engine_log()
    jp, pp, sp = str(tmp_path / "T_1_vmaf.json"), str(tmp_path / "T_1_psnr.txt"), str(tmp_path / "T_1_ssim.txt")
    report.write_libvmaf_json(jp, frames, pooled, 1234.5)
    report.write_ffmpeg_psnr_stats(pp, [{"mse": [4.0, 2.0, 2.5], "areas": [640 * 360, 320 * 180, 320 * 180]}] * 5, 8)
    report.write_ffmpeg_ssim_stats(sp, [{"ssim": [0.97, 0.98, 0.985], "weights": [4, 1, 1]}] * 5)
    az = ref_module.VMAFAnalyzer()
    res = az._parse_vmaf_results(jp, pp, sp, str(tmp_path / "dis.y4m"), str(tmp_path / "ref.y4m"))
    assert res is not None, az.error_occurred.sent
    raw = json.load(open(jp))
    # what pqa2_b200.vmaf_analyzer returns for the same log (vmaf_analyzer.py::_analyze, reference keys :919-932)
    assert res["vmaf_score"] == raw["pooled_metrics"]["vmaf"]["mean"] == float("%.6f" % pooled["vmaf"]["mean"])
    assert res["psnr_score"] == "T_1_psnr.txt" and res["ssim_score"] == "T_1_ssim.txt"
    assert res["json_path"] == jp and res["psnr_log"] == pp and res["ssim_log"] == sp
    assert res["reference_video"] == "ref.y4m" and res["distorted_video"] == "dis.y4m"
    assert res["raw_results"] == raw
    assert res["model"] == raw["version"]                         # no "model" key in a libvmaf log: the version string (:834-838)
    assert set(res) == {"vmaf_score", "psnr_score", "ssim_score", "json_path", "psnr_log", "ssim_log", "reference_video",
                        "distorted_video", "raw_results", "model", "width", "height"}
    assert az.analysis_progress.sent[-1] == (100,) and az.analysis_complete.sent[-1][0] is res
    # the fallback branch (:666-690) averages frames[].metrics when pooled_metrics is missing
    del raw["pooled_metrics"]
    json.dump(raw, open(jp, "w"))
    res2 = az._parse_vmaf_results(jp, pp, sp, "dis.y4m", "ref.y4m")
    assert abs(res2["vmaf_score"] - np.mean([f["metrics"]["vmaf"] for f in raw["frames"]])) < 1e-12


def test_csv_export_header_matches_libvmaf_with_psnr_and_ssim(ref_module, tmp_path):
    """results_tab.py:3009-3026: header = 'Frame Number' + sorted(first frame's metric keys).  With `psnr=1:ssim=1`
    (FFmpeg maps psnr=1 to the psnr extractor with enable_chroma=false) libvmaf logs psnr_y and float_ssim next to the
    model's features; so must this engine.  The extractor's chroma columns appear with EngineOptions.psnr_chroma."""
    from pqa2_b200 import report
    frames, pooled = _engine_log()
    assert "psnr_cb" not in frames[0]["metrics"]
    assert {"psnr_cb", "psnr_cr"} <= set(_engine_log(chroma=True)[0][0]["metrics"])
    jp = str(tmp_path / "T_vmaf.json")
    report.write_libvmaf_json(jp, frames, pooled, 100.0)
    raw = json.load(open(jp))
    # the lines of results_tab.py:3006-3013, restated on the parsed log
    first = raw["frames"][0]
    available = sorted(list(first.get("metrics", {}).keys()))
    want = sorted(["integer_adm2", "integer_adm_scale0", "integer_adm_scale1", "integer_adm_scale2", "integer_adm_scale3",
                   "integer_motion", "integer_motion2", "integer_vif_scale0", "integer_vif_scale1", "integer_vif_scale2",
                   "integer_vif_scale3", "psnr_y", "float_ssim", "vmaf"])
    assert available == want
    for fr in raw["frames"]:
        for k in available:
            assert isinstance(fr["metrics"][k], float) and f"{fr['metrics'][k]:.4f}"       # :3022-3024 formats every value
    # our own CSV writer produces the same header row
    csv_path = str(tmp_path / "T_data.csv")
    report.write_result_csv(csv_path, "T", {"frames": frames, "pooled_metrics": pooled, "model": "vmaf_v0.6.1", "n_frames": 5},
                            "ref.y4m", "dis.y4m")
    text = open(csv_path).read()
    assert "Frame Number," + ",".join(want) in text
