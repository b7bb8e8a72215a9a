"""The libvmaf filter option string of the reference, both ways.

* ``build_libvmaf_filter`` restates the option builder of ``app/vmaf_analyzer.py:373-406`` (same items, same
  order, joined with ``:``) -- kept so logs and the ``metadata.vmaf_options`` field stay comparable.
* ``parse_libvmaf_filter`` turns such a string -- the current one, or the legacy call site's
  ``libvmaf=log_fmt=json:log_path=...:psnr=1:ssim=1[:model_path=...]`` (``app/ui/tabs/results_tab.py:346-350``)
  -- into what the B200 engine needs: model, EngineOptions, log path, headline pool method.
"""
from __future__ import annotations

from . import engine


def build_libvmaf_filter(json_path: str, model: str = "vmaf_v0.6.1", threads: int = 4, feature_subsample: int = 1,
                         pool_method: str = "mean", enable_motion_score: bool = False,
                         enable_temporal_features: bool = False) -> str:
    sep = any(c in model for c in ("/", "\\"))
    opts = [f"log_path={json_path}", "log_fmt=json",
            f"model=path={model}" if sep else f"model=version={model}",
            f"n_threads={threads}", f"n_subsample={feature_subsample}"]
    if pool_method != "mean":
        opts += [f"pool={pool_method}", "psnr=1", "ssim=1"]
    if enable_motion_score:
        opts.append("feature=name=motion:enable=1")
    if enable_temporal_features:
        opts += [f"feature=name={n}:enable=1" for n in ("vif_scale0", "vif_scale1", "vif_scale2", "vif_scale3", "adm2", "motion")]
        if enable_motion_score:
            opts.append("feature=name=motion:enable=1")
    return "libvmaf=" + ":".join(opts)


def _split(s: str) -> list:
    """Split on ':' but keep 'feature=name=x:enable=1' and Windows drive letters ('C:/...') together."""
    parts, cur = [], ""
    for tok in s.split(":"):
        if cur and (tok.startswith("enable=") or (len(cur.rsplit("=", 1)[-1]) == 1 and tok[:1] in "/\\")):
            cur += ":" + tok
        else:
            if cur:
                parts.append(cur)
            cur = tok
    if cur:
        parts.append(cur)
    return parts


def parse_libvmaf_filter(s: str) -> dict:
    """-> {"model": str, "log_path": str|None, "log_fmt": str, "pool": str, "options": EngineOptions,
    "n_threads": int, "features": [names]}; unknown keys are kept under "extra"."""
    if s.startswith("libvmaf="):
        s = s[len("libvmaf="):]
    out = {"model": "vmaf_v0.6.1", "log_path": None, "log_fmt": "xml", "pool": "mean", "n_threads": 0, "features": [],
           "extra": {}}
    o = engine.EngineOptions()
    for item in _split(s):
        if not item:
            continue
        k, _, v = item.partition("=")
        if k == "log_path":
            out["log_path"] = v
        elif k == "log_fmt":
            out["log_fmt"] = v
        elif k == "model":
            kk, _, vv = v.partition("=")
            out["model"] = vv if kk in ("version", "path") else v
        elif k == "model_path":                                     # legacy spelling (libvmaf < 2)
            out["model"] = v
        elif k == "n_threads":
            out["n_threads"] = int(v)
        elif k == "n_subsample":
            o.n_subsample = max(1, int(v))
        elif k == "pool":
            out["pool"] = v
        elif k == "psnr":
            o.psnr = v not in ("0", "false")
        elif k == "ssim":
            o.ssim = v not in ("0", "false")
        elif k == "ms_ssim":
            o.ms_ssim = v not in ("0", "false")
        elif k == "feature":
            name = v.partition("=")[2].partition(":")[0]
            out["features"].append(name)
            if name == "psnr":
                o.psnr = True
            elif name == "float_ssim":
                o.ssim = True
            elif name == "float_ms_ssim":
                o.ms_ssim = True
        else:
            out["extra"][k] = v
    out["options"] = o
    return out
