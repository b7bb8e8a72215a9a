"""Raw-video ingest for the engine: Y4M and headerless planar YUV readers/writers.

The reference hands ffmpeg two container files (``app/vmaf_analyzer.py:415-416``) and lets it
decode; libvmaf then sees planar pictures, 8-bit as u8 and >8-bit as little-endian u16
(SURVEY.md Appendix A.1).  This engine consumes those planar pictures directly; compressed inputs (the
reference's aligned MP4s) are decoded on the host by the libavcodec inside the cv2 wheel (``avdec``)."""
from __future__ import annotations

import os
import re
from dataclasses import dataclass

import numpy as np


@dataclass
class ClipInfo:
    path: str
    width: int
    height: int
    bpc: int
    chroma: int            # 420 / 422 / 444 / 400
    fps_num: int = 30
    fps_den: int = 1
    nb_frames: int = 0
    header_bytes: int = 0
    frame_header_bytes: int = 0
    decoder: str = "raw"          # "raw" (y4m / planar yuv), "av" (container through the cv2 wheel's libavcodec: all planes)
                                  # or "cv2" (container through cv2.VideoCapture: luma only)
    codec_name: str = "rawvideo"  # ffprobe's codec_name (app/vmaf_analyzer.py:221-231); containers report their stream's codec

    @property
    def fps(self) -> float:
        return self.fps_num / self.fps_den if self.fps_den else 0.0

    @property
    def pix_fmt(self) -> str:
        base = {420: "yuv420p", 422: "yuv422p", 444: "yuv444p", 400: "gray"}[self.chroma]
        return base if self.bpc == 8 else f"{base}{self.bpc}le"

    def plane_shapes(self):
        w, h = self.width, self.height
        if self.chroma == 400:
            return [(h, w)]
        cw = (w + 1) // 2 if self.chroma in (420, 422) else w
        ch = (h + 1) // 2 if self.chroma == 420 else h
        return [(h, w), (ch, cw), (ch, cw)]

    @property
    def frame_bytes(self) -> int:
        bps = 1 if self.bpc == 8 else 2
        return sum(a * b for a, b in self.plane_shapes()) * bps


_Y4M_CHROMA = {
    "420": (420, 8), "420jpeg": (420, 8), "420mpeg2": (420, 8), "420paldv": (420, 8),
    "422": (422, 8), "444": (444, 8), "mono": (400, 8),
    "420p10": (420, 10), "422p10": (422, 10), "444p10": (444, 10), "mono10": (400, 10),
    "420p12": (420, 12), "422p12": (422, 12), "444p12": (444, 12), "mono12": (400, 12),
    "420p16": (420, 16), "422p16": (422, 16), "444p16": (444, 16),
}


def probe(path: str, width: int | None = None, height: int | None = None, pix_fmt: str | None = None,
          fps: float | None = None) -> ClipInfo:
    """Metadata of a raw clip.  ``.y4m`` is self-describing; ``.yuv`` needs width/height/pix_fmt or a
    ``_<W>x<H>[_<fps>][_<pix_fmt>]`` hint in the file name."""
    size = os.path.getsize(path)
    with open(path, "rb") as f:
        head = f.read(10)
        if head.startswith(b"YUV4MPEG2"):
            f.seek(0)
            line = f.readline(4096)
            toks = line.decode("ascii", "replace").strip().split(" ")
            w = h = 0
            chroma, bpc = 420, 8
            fn, fd = 30, 1
            for t in toks[1:]:
                if not t:
                    continue
                if t[0] == "W":
                    w = int(t[1:])
                elif t[0] == "H":
                    h = int(t[1:])
                elif t[0] == "F":
                    a, _, b = t[1:].partition(":")
                    fn, fd = int(a), int(b or 1)
                elif t[0] == "C":
                    if t[1:] not in _Y4M_CHROMA:
                        raise ValueError(f"{path}: unsupported Y4M chroma tag {t}")
                    chroma, bpc = _Y4M_CHROMA[t[1:]]
            if w <= 0 or h <= 0:
                raise ValueError(f"{path}: Y4M header lacks W/H")
            info = ClipInfo(path, w, h, bpc, chroma, fn, fd, header_bytes=len(line))
            # frame headers are "FRAME\n" unless they carry parameters; measure the first one
            fh = f.readline(256)
            info.frame_header_bytes = len(fh) if fh.startswith(b"FRAME") else 6
            per = info.frame_header_bytes + info.frame_bytes
            info.nb_frames = (size - info.header_bytes) // per
            return info
    name = os.path.basename(path)
    if os.path.splitext(name)[1].lower() in _CONTAINER_EXT:
        return _probe_container(path)
    if width is None or height is None:
        m = re.search(r"(\d{2,5})x(\d{2,5})", name)
        if not m:
            raise ValueError(f"{path}: raw YUV needs width/height (or a _WxH hint in the name)")
        width, height = int(m.group(1)), int(m.group(2))
    if pix_fmt is None:
        m = re.search(r"(yuv4(?:20|22|44)p(?:1[026](?:le)?)?|gray(?:1[026]le)?)", name)
        pix_fmt = m.group(1) if m else "yuv420p"
    m = re.match(r"(?:yuv(4\d\d)p|(gray))(\d\d)?(?:le)?$", pix_fmt)
    if not m:
        raise ValueError(f"unsupported pix_fmt {pix_fmt}")
    chroma = 400 if m.group(2) else int(m.group(1))
    bpc = int(m.group(3)) if m.group(3) else 8
    fn, fd = (int(round((fps or 30.0) * 1000)), 1000)
    info = ClipInfo(path, width, height, bpc, chroma, fn, fd)
    info.nb_frames = size // info.frame_bytes
    return info


_CONTAINER_EXT = {".mp4", ".mov", ".mkv", ".avi", ".m4v", ".webm", ".ts"}


def _probe_container(path: str) -> ClipInfo:
    """Compressed clips (the reference's aligned H.264 MP4s, app/bookend_alignment.py:526-536).  Preferred decoder:
    ``avdec`` -- the libavformat / libavcodec that ship inside the cv2 wheel, driven directly, which yields Y, U and V as
    the reference's ffmpeg child would hand them to libvmaf and to the psnr / ssim filters.  Without it: cv2's
    VideoCapture, which with CAP_PROP_CONVERT_RGB off hands back the decoder's luma plane untouched and nothing else, so
    such clips are described as 8-bit luma-only (enough for VMAF, psnr_y, float_ssim, float_ms_ssim; the all-plane
    stats files are skipped)."""
    try:
        import cv2
    except Exception as e:                                  # noqa: BLE001
        raise ValueError(f"{path}: container input needs cv2 (not installed): {e}")
    cap = cv2.VideoCapture(path)
    try:
        if not cap.isOpened():
            raise ValueError(f"{path}: cv2 cannot open this file")
        w, h = int(cap.get(cv2.CAP_PROP_FRAME_WIDTH)), int(cap.get(cv2.CAP_PROP_FRAME_HEIGHT))
        n = int(cap.get(cv2.CAP_PROP_FRAME_COUNT))          # the container's own count, or duration x rate: an estimate
        fps = float(cap.get(cv2.CAP_PROP_FPS)) or 30.0
        cc = int(cap.get(cv2.CAP_PROP_FOURCC))
    finally:
        cap.release()
    if w <= 0 or h <= 0:
        raise ValueError(f"{path}: no video stream")
    tag = "".join(chr((cc >> (8 * k)) & 0xFF) for k in range(4)).strip("\0 ").lower()
    from . import avdec
    if avdec.available():
        try:
            with avdec.AvDecoder(path, threads=1) as d:
                return ClipInfo(path, d.width, d.height, d.bpc, d.chroma, d.fps_num, d.fps_den, nb_frames=n, decoder="av",
                                codec_name=d.codec_name)
        except (RuntimeError, ValueError):                  # layout check failed / unsupported pixel format: luma via cv2
            pass
    return ClipInfo(path, w, h, 8, 400, int(round(fps * 1000)), 1000, nb_frames=n, decoder="cv2",
                    codec_name=_FOURCC_CODEC.get(tag, tag or "unknown"))


# container fourcc -> the codec_name ffprobe prints for it
_FOURCC_CODEC = {"avc1": "h264", "h264": "h264", "x264": "h264", "hev1": "hevc", "hvc1": "hevc", "hevc": "hevc",
                 "mp4v": "mpeg4", "fmp4": "mpeg4", "xvid": "mpeg4", "vp09": "vp9", "vp90": "vp9", "vp80": "vp8",
                 "av01": "av1", "mjpg": "mjpeg", "apch": "prores", "apcn": "prores", "ffv1": "ffv1"}


class EndOfClip(EOFError):
    """The decoder ran out of frames before the container's (estimated) frame count: the clip ends here, as it would
    for ffmpeg + libvmaf, which stop at the shorter input."""


class _Cv2Reader:
    """Luma planes of a container file, decoded by cv2.  Strictly sequential: CAP_PROP_POS_FRAMES is not frame-exact
    for H.264 with B-frames / VFR streams, so a frame further ahead is reached by decoding and discarding, and a
    frame behind the cursor by reopening the file -- never by seeking."""

    def __init__(self, info: ClipInfo):
        import cv2
        self.info, self._cv2 = info, cv2
        try:                                               # cv2 warns on every frame that yuv420p is handed back raw
            cv2.utils.logging.setLogLevel(cv2.utils.logging.LOG_LEVEL_ERROR)
        except Exception:                                  # noqa: BLE001
            pass
        self._cap = None
        self._open()

    def _open(self):
        if self._cap is not None:
            self._cap.release()
        self._cap = self._cv2.VideoCapture(self.info.path)
        self._cap.set(self._cv2.CAP_PROP_CONVERT_RGB, 0)
        self._next = 0

    def close(self):
        self._cap.release()

    def alloc_planes(self, pinned: bool = True):
        if pinned:
            from .extractor import pinned_empty
            return [pinned_empty((self.info.height, self.info.width), np.uint8)]
        return [np.empty((self.info.height, self.info.width), np.uint8)]

    def read_into(self, i: int, planes, luma_only: bool = False) -> None:
        if i < self._next:
            self._open()
        while self._next < i:                               # decode and discard up to the requested frame
            if not self._cap.grab():
                raise EndOfClip(f"{self.info.path}: stream ends at frame {self._next}")
            self._next += 1
        ok, fr = self._cap.read()
        if not ok or fr is None:
            if i == 0:
                raise EOFError(f"{self.info.path}: cannot decode the first frame")
            raise EndOfClip(f"{self.info.path}: stream ends at frame {i}")
        self._next = i + 1
        h, w = self.info.height, self.info.width
        if fr.ndim == 3:
            # the backend ignored CONVERT_RGB=0 and converted to BGR: recomputing Y from it is a range / matrix round
            # trip, not the decoder's luma, and would change scores silently
            raise ValueError(f"{self.info.path}: this cv2 build does not hand back the decoder's luma plane")
        y = fr.reshape(-1, w)[:h]                           # (h, w) luma; some builds return the whole I420 buffer (3h/2, w)
        planes[0][...] = y


class _AvReader:
    """All planes of a container file through avdec (libavformat / libavcodec).  Strictly sequential, like _Cv2Reader:
    a frame further ahead is reached by decoding and discarding, a frame behind the cursor by reopening the file."""

    def __init__(self, info: ClipInfo):
        self.info = info
        self._dec = None
        self._dtype = np.uint8 if info.bpc == 8 else np.dtype("<u2")
        self._open()

    def _open(self):
        from . import avdec
        if self._dec is not None:
            self._dec.close()
        self._dec = avdec.AvDecoder(self.info.path)
        d, inf = self._dec, self.info
        if (d.width, d.height, d.bpc, d.chroma) != (inf.width, inf.height, inf.bpc, inf.chroma):
            raise ValueError(f"{inf.path}: stream changed since it was probed")
        self._next = 0

    def close(self):
        if self._dec is not None:
            self._dec.close()
            self._dec = None

    def alloc_planes(self, pinned: bool = True):
        if pinned:
            from .extractor import pinned_empty
            return [pinned_empty(s, self._dtype) for s in self.info.plane_shapes()]
        return [np.empty(s, self._dtype) for s in self.info.plane_shapes()]

    def read_into(self, i: int, planes, luma_only: bool = False) -> None:
        if i < self._next:
            self._open()
        while self._next < i:
            if not self._dec.next():
                raise EndOfClip(f"{self.info.path}: stream ends at frame {self._next}")
            self._next += 1
        if not self._dec.next(planes, luma_only):
            if i == 0:
                raise EOFError(f"{self.info.path}: cannot decode the first frame")
            raise EndOfClip(f"{self.info.path}: stream ends at frame {i}")
        self._next = i + 1


class ClipReader:
    """Sequential / random access reader that fills caller-provided (pinned) plane arrays."""

    def __new__(cls, info: ClipInfo):
        dec = getattr(info, "decoder", "raw")
        if dec == "cv2":
            return _Cv2Reader(info)
        if dec == "av":
            return _AvReader(info)
        return super().__new__(cls)

    def __init__(self, info: ClipInfo):
        self.info = info
        self._f = open(info.path, "rb", buffering=0)
        self._dtype = np.uint8 if info.bpc == 8 else np.dtype("<u2")

    def close(self):
        self._f.close()

    def frame_offset(self, i: int) -> int:
        inf = self.info
        return inf.header_bytes + i * (inf.frame_header_bytes + inf.frame_bytes) + inf.frame_header_bytes

    def alloc_planes(self, pinned: bool = True):
        if pinned:
            from .extractor import pinned_empty
            return [pinned_empty(s, self._dtype) for s in self.info.plane_shapes()]
        return [np.empty(s, self._dtype) for s in self.info.plane_shapes()]

    def read_into(self, i: int, planes, luma_only: bool = False) -> None:
        flat = getattr(planes, "flat", None)
        if flat is not None and not luma_only and len(planes) == len(self.info.plane_shapes()) and \
                flat.nbytes == self.info.frame_bytes:
            # the planes are views of one buffer laid out like the file's frame: one read for Y, U and V
            n = os.preadv(self._f.fileno(), [memoryview(flat)], self.frame_offset(i))
            if n != flat.nbytes:
                raise EOFError(f"{self.info.path}: short read at frame {i}")
            return
        self._f.seek(self.frame_offset(i))
        for k, p in enumerate(planes):
            if luma_only and k > 0:
                break
            mv = memoryview(p.reshape(-1).view(np.uint8))
            n = self._f.readinto(mv)
            if n != len(mv):
                raise EOFError(f"{self.info.path}: short read at frame {i}")


class MappedClip:
    """A raw clip mapped into memory and registered with CUDA (``bv_host_register``): frame planes are numpy views
    of the page cache, and ``bv_submit`` DMAs them to the GPU without a CPU copy.  ``MappedClip.open`` returns None
    when the file cannot be mapped or the platform refuses the registration (the caller then reads through a pinned
    ring, ``ClipReader``)."""

    def __init__(self, info: ClipInfo, mm, base_addr: int, registered_len: int):
        self.info, self._mm, self._addr, self._len = info, mm, base_addr, registered_len
        self._dtype = np.dtype(np.uint8) if info.bpc == 8 else np.dtype("<u2")
        self._buf = np.frombuffer(mm, dtype=np.uint8)
        self._shapes = info.plane_shapes()

    @classmethod
    def open(cls, info: ClipInfo, min_bytes: int = 1 << 20):
        import ctypes
        import mmap
        from . import _lib as L
        if getattr(info, "decoder", "raw") != "raw" or info.nb_frames < 1:
            return None
        size = os.path.getsize(info.path)
        if size < min_bytes:
            return None
        lib = L.load()
        # a read-only shared mapping with a read-only registration first; where that is refused, a writable shared
        # mapping (never written to) with a plain registration
        for mode, access, ro in (("rb", mmap.ACCESS_READ, 1), ("r+b", mmap.ACCESS_WRITE, 0)):
            try:
                with open(info.path, mode) as f:
                    mm = mmap.mmap(f.fileno(), 0, access=access)
            except (OSError, ValueError):
                continue
            arr = np.frombuffer(mm, dtype=np.uint8)
            addr = arr.ctypes.data
            del arr
            if lib.bv_host_register(ctypes.c_void_p(addr), size, ro) == 0:
                m = cls(info, mm, addr, size)
                m.mode = "read-only mapping" if ro else "writable shared mapping"
                return m
            try:
                mm.close()
            except BufferError:
                pass
        return None

    def planes(self, i: int, luma_only: bool = False):
        inf = self.info
        off = inf.header_bytes + i * (inf.frame_header_bytes + inf.frame_bytes) + inf.frame_header_bytes
        if i < 0 or off + inf.frame_bytes > self._len:
            raise EOFError(f"{inf.path}: frame {i} is past the end of the file")
        out = []
        bps = self._dtype.itemsize
        for (h, w) in self._shapes[: 1 if luma_only else len(self._shapes)]:
            n = h * w * bps
            a = self._buf[off:off + n]
            out.append(a.view(self._dtype).reshape(h, w))       # 16-bit planes may sit at odd addresses: the copy engine does not care
            off += n
        return out

    def close(self):
        if self._mm is None:
            return
        import ctypes
        from . import _lib as L
        L.load().bv_host_unregister(ctypes.c_void_p(self._addr))
        self._buf = None
        try:
            self._mm.close()
        except BufferError:            # a view is still alive somewhere; the mapping goes with it
            pass
        self._mm = None


def write_y4m(path: str, frames, width: int, height: int, bpc: int = 8, fps=(30, 1), chroma: int = 420) -> None:
    """frames: iterable of [Y, U, V] arrays.  Header as ffmpeg expects."""
    tag = {(420, 8): "420jpeg", (422, 8): "422", (444, 8): "444", (400, 8): "mono"}.get((chroma, bpc))
    if tag is None:
        tag = {420: "420", 422: "422", 444: "444", 400: "mono"}[chroma] + ("p" if chroma != 400 else "") + str(bpc)
    extra = f" XYSCSS={chroma}P{bpc}" if bpc > 8 and chroma != 400 else ""
    with open(path, "wb") as f:
        f.write(f"YUV4MPEG2 W{width} H{height} F{fps[0]}:{fps[1]} Ip A1:1 C{tag}{extra}\n".encode())
        for planes in frames:
            f.write(b"FRAME\n")
            for p in planes:
                a = np.ascontiguousarray(p)
                f.write(a.astype("<u2").tobytes() if bpc > 8 else a.tobytes())


def write_raw(path: str, frames, bpc: int = 8) -> None:
    with open(path, "wb") as f:
        for planes in frames:
            for p in planes:
                a = np.ascontiguousarray(p)
                f.write(a.astype("<u2").tobytes() if bpc > 8 else a.tobytes())
