"""VMAF model files: loader for libvmaf's JSON models and this repo's packed form, plus the SVR
fusion wrapper over the C ABI (``bv_model_create`` / ``bv_predict`` / ``bv_predict_device``).

Schema followed: reference ``models/vmaf_v0.6.1.json:1-69`` (``model_dict``: libsvm text under
``model``, ``feature_names``, ``slopes``/``intercepts`` with index 0 = score, ``score_clip``,
``score_transform``, ``feature_opts_dicts``), the NEG variant ``models/vmaf_v0.6.1neg.json:34-51`` and
the bootstrap collection ``models/vmaf_b_v0.6.3.json`` (keys "0".."20").  The reference only lists these
files in a dropdown (``app/ui/tabs/analysis_tab.py:1005-1048``) and hands the stem to libvmaf
(``app/vmaf_analyzer.py:377``); parsing them is libvmaf's ``read_json_model.c``, restated here."""
from __future__ import annotations

import ctypes as C
import json
import os
import threading
from dataclasses import dataclass, field

import numpy as np

from . import _lib as L

MODELS_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "models")
_HANDLE_LOCK = threading.Lock()          # SvrModel.handle(): one bv_model per SvrModel even when worker threads race


@dataclass
class SvrModel:
    """One nu-SVR (RBF) with libvmaf's linear_rescale normalisation."""
    name: str
    feature_names: list            # e.g. VMAF_integer_feature_adm2_score
    sv: np.ndarray                 # [n_sv, n_feat] dense (sparse indices missing in the file are 0)
    coef: np.ndarray               # [n_sv]
    gamma: float
    rho: float
    slopes: np.ndarray             # [n_feat + 1], index 0 = score
    intercepts: np.ndarray
    score_clip: tuple | None
    transform: dict                # p0, p1, p2, out_lte_in, out_gte_in (absent keys omitted)
    feature_opts: list = field(default_factory=list)
    _handle: int | None = None

    # ---- derived -----------------------------------------------------------------------------
    @property
    def is_float(self) -> bool:
        return not any("integer" in n for n in self.feature_names)

    @property
    def metric_keys(self) -> list:
        """libvmaf log names of the model's inputs, in model order."""
        out = []
        for n in self.feature_names:
            k = n
            if k.startswith("VMAF_integer_feature_"):
                k = "integer_" + k[len("VMAF_integer_feature_"):]
            elif k.startswith("VMAF_feature_"):
                k = k[len("VMAF_feature_"):]
            if k.endswith("_score"):
                k = k[: -len("_score")]
            out.append(k)
        return out

    def opt(self, key: str, default: float) -> float:
        for d in self.feature_opts:
            if key in d:
                return float(d[key])
        return default

    # ---- C ABI -------------------------------------------------------------------------------
    def handle(self):
        if self._handle is not None:
            return self._handle
        with _HANDLE_LOCK:
            if self._handle is not None:
                return self._handle
            lib = L.load()
            pd = C.POINTER(C.c_double)
            sv = np.ascontiguousarray(self.sv, np.float64)
            coef = np.ascontiguousarray(self.coef, np.float64)
            sl = np.ascontiguousarray(self.slopes, np.float64)
            ic = np.ascontiguousarray(self.intercepts, np.float64)
            clip = np.array(self.score_clip if self.score_clip else (0.0, 0.0), np.float64)
            tp = np.array([self.transform.get("p0", 0.0), self.transform.get("p1", 0.0),
                           self.transform.get("p2", 0.0)], np.float64)
            tf = 0
            for bit, key in enumerate(("p0", "p1", "p2")):
                if key in self.transform:
                    tf |= 1 << bit
            if self.transform.get("out_lte_in"):
                tf |= 8
            if self.transform.get("out_gte_in"):
                tf |= 16
            h = lib.bv_model_create(sv.shape[1], sv.shape[0], sv.ctypes.data_as(pd), coef.ctypes.data_as(pd),
                                    float(self.gamma), float(self.rho), sl.ctypes.data_as(pd), ic.ctypes.data_as(pd),
                                    clip.ctypes.data_as(pd), 1 if self.score_clip else 0, tp.ctypes.data_as(pd), tf)
            if not h:
                raise RuntimeError("bv_model_create failed")
            self._handle = h
        return self._handle

    def predict(self, feats: np.ndarray, enable_transform: bool = False, disable_clip: bool = False,
                device: int | None = None) -> np.ndarray:
        """feats: [n, n_feat] in model order.  device=None -> host libsvm-order evaluation in the C
        library; device=k -> the svr_predict CUDA kernel on GPU k."""
        lib = L.load()
        feats = np.ascontiguousarray(feats, np.float64).reshape(-1, self.sv.shape[1])
        out = np.empty(feats.shape[0], np.float64)
        flags = (L.MODEL_ENABLE_TRANSFORM if enable_transform else 0) | (L.MODEL_DISABLE_CLIP if disable_clip else 0)
        pd = C.POINTER(C.c_double)
        if device is None:
            rc = lib.bv_predict(self.handle(), feats.ctypes.data_as(pd), feats.shape[0], flags, out.ctypes.data_as(pd))
        else:
            rc = lib.bv_predict_device(self.handle(), device, feats.ctypes.data_as(pd), feats.shape[0], flags,
                                       out.ctypes.data_as(pd))
        if rc != 0:
            raise RuntimeError(f"bv_predict failed ({rc})")
        return out

    def __del__(self):
        try:
            if self._handle is not None:
                L.load(build_if_missing=False).bv_model_free(self._handle)
                self._handle = None
        except Exception:
            pass


@dataclass
class VmafModel:
    """A model file: one SVR, or a bootstrap collection (main model + n bootstrap models)."""
    name: str
    main: SvrModel
    bootstrap: list = field(default_factory=list)

    @property
    def is_float(self) -> bool:
        return self.main.is_float

    @property
    def vif_enhn_gain_limit(self) -> float:
        return self.main.opt("vif_enhn_gain_limit", 100.0)

    @property
    def adm_enhn_gain_limit(self) -> float:
        return self.main.opt("adm_enhn_gain_limit", 100.0)


def parse_libsvm_text(text: str, n_feat: int):
    """libsvm model text (svm.cpp svm_load_model): header lines, then `coef idx:val ...` rows."""
    gamma = rho = None
    total_sv = None
    lines = text.split("\n")
    i = 0
    while i < len(lines):
        ln = lines[i].strip()
        i += 1
        if not ln:
            continue
        if ln == "SV":
            break
        key, _, val = ln.partition(" ")
        if key == "svm_type" and val != "nu_svr":
            raise ValueError(f"unsupported svm_type {val}")
        if key == "kernel_type" and val != "rbf":
            raise ValueError(f"unsupported kernel_type {val}")
        if key == "gamma":
            gamma = float(val)
        if key == "rho":
            rho = float(val)
        if key == "total_sv":
            total_sv = int(val)
    coef, sv = [], []
    for ln in lines[i:]:
        ln = ln.strip()
        if not ln:
            continue
        parts = ln.split()
        coef.append(float(parts[0]))
        row = [0.0] * n_feat
        for tok in parts[1:]:
            idx, _, v = tok.partition(":")
            k = int(idx)
            if 1 <= k <= n_feat:
                row[k - 1] = float(v)
        sv.append(row)
    if gamma is None or rho is None:
        raise ValueError("libsvm text lacks gamma/rho")
    if total_sv is not None and total_sv != len(sv):
        raise ValueError(f"total_sv {total_sv} != {len(sv)} SV rows")
    return np.array(sv, np.float64).reshape(len(sv), n_feat), np.array(coef, np.float64), gamma, rho


def _truthy(v) -> bool:
    return v is True or (isinstance(v, str) and v.lower() == "true")


def _svr_from_model_dict(name: str, md: dict) -> SvrModel:
    names = list(md["feature_names"])
    sv, coef, gamma, rho = parse_libsvm_text(md["model"], len(names))
    if md.get("norm_type", "linear_rescale") != "linear_rescale":
        raise ValueError(f"unsupported norm_type {md.get('norm_type')}")
    tr = {}
    st = md.get("score_transform") or {}
    for k in ("p0", "p1", "p2"):
        if k in st:
            tr[k] = float(st[k])
    for k in ("out_lte_in", "out_gte_in"):
        if _truthy(st.get(k)):
            tr[k] = True
    clip = tuple(float(x) for x in md["score_clip"]) if md.get("score_clip") else None
    return SvrModel(name=name, feature_names=names, sv=sv, coef=coef, gamma=gamma, rho=rho,
                    slopes=np.array(md["slopes"], np.float64), intercepts=np.array(md["intercepts"], np.float64),
                    score_clip=clip, transform=tr, feature_opts=list(md.get("feature_opts_dicts") or []))


def _svr_from_packed(name: str, d: dict) -> SvrModel:
    return SvrModel(name=name, feature_names=list(d["feature_names"]),
                    sv=np.array(d["sv"], np.float64).reshape(len(d["coef"]), len(d["feature_names"])),
                    coef=np.array(d["coef"], np.float64), gamma=float(d["gamma"]), rho=float(d["rho"]),
                    slopes=np.array(d["slopes"], np.float64), intercepts=np.array(d["intercepts"], np.float64),
                    score_clip=tuple(d["score_clip"]) if d.get("score_clip") else None,
                    transform=dict(d.get("transform") or {}), feature_opts=list(d.get("feature_opts") or []))


def pack(model: VmafModel) -> dict:
    """Packed, dense form written by tools/pack_models.py (hex floats keep every bit)."""
    def one(m: SvrModel) -> dict:
        return {"feature_names": m.feature_names, "gamma": m.gamma, "rho": m.rho,
                "coef": [float(x) for x in m.coef], "sv": [float(x) for x in m.sv.ravel()],
                "slopes": [float(x) for x in m.slopes], "intercepts": [float(x) for x in m.intercepts],
                "score_clip": list(m.score_clip) if m.score_clip else None, "transform": m.transform,
                "feature_opts": m.feature_opts}
    return {"format": "b200vmaf-packed-1", "name": model.name, "main": one(model.main),
            "bootstrap": [one(b) for b in model.bootstrap]}


def load_model_file(path: str) -> VmafModel:
    with open(path, "r") as f:
        data = json.load(f)
    name = os.path.basename(path)
    for suf in (".bvm.json", ".json"):
        if name.endswith(suf):
            name = name[: -len(suf)]
            break
    if data.get("format") == "b200vmaf-packed-1":
        return VmafModel(name=data.get("name", name), main=_svr_from_packed(name, data["main"]),
                         bootstrap=[_svr_from_packed(f"{name}#{i + 1}", b) for i, b in enumerate(data["bootstrap"])])
    if "model_dict" in data:
        return VmafModel(name=name, main=_svr_from_model_dict(name, data["model_dict"]))
    if "0" in data and "model_dict" in data["0"]:
        keys = sorted((k for k in data if k.isdigit()), key=int)
        subs = [_svr_from_model_dict(f"{name}#{k}", data[k]["model_dict"]) for k in keys]
        return VmafModel(name=name, main=subs[0], bootstrap=subs[1:])
    raise ValueError(f"{path}: not a libvmaf JSON model")


def resolve_model(model: str | None) -> VmafModel:
    """Reference semantics (app/vmaf_analyzer.py:328-331, :377): None -> vmaf_v0.6.1; a bare stem is a
    built-in version; anything with a path separator is a file path."""
    if model is None:
        model = "vmaf_v0.6.1"
    if model.startswith("path="):
        model = model[5:]
    if model.startswith("version="):
        model = model[8:]
    if any(sep in model for sep in ("/", "\\")) or os.path.isfile(model):
        return load_model_file(model)
    stem = model[:-5] if model.endswith(".json") else model
    for cand in (os.path.join(MODELS_DIR, stem + ".bvm.json"), os.path.join(MODELS_DIR, stem + ".json")):
        if os.path.isfile(cand):
            return load_model_file(cand)
    raise FileNotFoundError(f"VMAF model '{model}' not found (built-in models live in {MODELS_DIR})")


def available_models() -> list:
    out = []
    for f in sorted(os.listdir(MODELS_DIR)):
        if f.endswith(".bvm.json"):
            out.append(f[: -len(".bvm.json")])
        elif f.endswith(".json"):
            out.append(f[: -len(".json")])
    return out
