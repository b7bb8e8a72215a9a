"""Container decode to planar Y, U, V through the libavformat / libavcodec that ship inside the cv2 wheel (ctypes).

What the reference's ffmpeg child does to ``-i distorted -i reference`` before libvmaf sees planar pictures
(``app/vmaf_analyzer.py:411-419``; its real inputs are the aligned H.264 MP4s of
``app/bookend_alignment.py:526-536``).  cv2's own ``VideoCapture`` only exposes the luma plane of a decoded frame
("yuv420p ... will be treated as 8UC1"), so the FFmpeg ``psnr`` / ``ssim`` passes over Y, Cb, Cr
(``app/vmaf_analyzer.py:996-1092``) could not run on compressed inputs; the libraries behind it export the whole public
FFmpeg API, and this module drives the few calls a sequential decoder needs.

Only PUBLIC structure prefixes are touched, each checked against a value known from elsewhere before it is trusted
(AVFormatContext.nb_streams / streams, AVStream.index / codecpar, AVCodecParameters.codec_type / codec_id, AVPacket up to
stream_index, AVFrame up to format): they have been stable since FFmpeg 5.1.  If the libraries are missing, of another
major version family, or a check fails, ``available()`` is False and the caller keeps the cv2 luma-only path."""
from __future__ import annotations

import ctypes as C
import glob
import os
import threading

import numpy as np

_AVERROR_EOF = -541478725          # -MKTAG('E','O','F',' ')
_AVERROR_EAGAIN = -11
_AVMEDIA_TYPE_VIDEO = 0

_lock = threading.Lock()
_libs = None                       # dict or False


class _Rational(C.Structure):
    _fields_ = [("num", C.c_int), ("den", C.c_int)]


class _PacketHead(C.Structure):    # AVPacket up to stream_index (allocated by av_packet_alloc, never by us)
    _fields_ = [("buf", C.c_void_p), ("pts", C.c_int64), ("dts", C.c_int64), ("data", C.c_void_p), ("size", C.c_int),
                ("stream_index", C.c_int)]


class _FrameHead(C.Structure):     # AVFrame up to format (allocated by av_frame_alloc)
    _fields_ = [("data", C.c_void_p * 8), ("linesize", C.c_int * 8), ("extended_data", C.c_void_p), ("width", C.c_int),
                ("height", C.c_int), ("nb_samples", C.c_int), ("format", C.c_int)]


# pix_fmt name -> (chroma, bits per component); full-range "yuvj" variants carry the same samples
_PIX = {"gray": (400, 8), "gray10le": (400, 10), "gray12le": (400, 12)}
for _c in (420, 422, 444):
    _PIX[f"yuv{_c}p"] = (_c, 8)
    _PIX[f"yuvj{_c}p"] = (_c, 8)
    for _b in (10, 12, 16):
        _PIX[f"yuv{_c}p{_b}le"] = (_c, _b)


def _load():
    global _libs
    with _lock:
        if _libs is not None:
            return _libs
        _libs = False
        try:
            import cv2
            base = os.path.join(os.path.dirname(os.path.dirname(cv2.__file__)), "opencv_python_headless.libs")
            if not os.path.isdir(base):
                base = os.path.join(os.path.dirname(os.path.dirname(cv2.__file__)), "opencv_python.libs")
            L = {}
            for n in ("avutil", "swresample", "avcodec", "avformat"):
                hits = sorted(glob.glob(os.path.join(base, f"lib{n}-*.so*")))
                if not hits:
                    return _libs
                L[n] = C.CDLL(hits[0], mode=C.RTLD_GLOBAL)
            if not (59 <= (L["avformat"].avformat_version() >> 16) <= 62 and 59 <= (L["avcodec"].avcodec_version() >> 16) <= 62):
                return _libs                           # structure prefixes verified for FFmpeg 5.1 .. 8.0 only
            vp, i = C.c_void_p, C.c_int
            F, K, U = L["avformat"], L["avcodec"], L["avutil"]
            F.avformat_open_input.argtypes = [C.POINTER(vp), C.c_char_p, vp, vp]
            F.avformat_find_stream_info.argtypes = [vp, vp]
            F.av_find_best_stream.argtypes = [vp, i, i, i, C.POINTER(vp), i]
            F.av_read_frame.argtypes = [vp, vp]
            F.avformat_close_input.argtypes = [C.POINTER(vp)]
            F.av_guess_frame_rate.argtypes = [vp, vp, vp]
            F.av_guess_frame_rate.restype = _Rational
            K.avcodec_alloc_context3.argtypes = [vp]
            K.avcodec_alloc_context3.restype = vp
            K.avcodec_parameters_to_context.argtypes = [vp, vp]
            K.avcodec_open2.argtypes = [vp, vp, vp]
            K.avcodec_send_packet.argtypes = [vp, vp]
            K.avcodec_receive_frame.argtypes = [vp, vp]
            K.avcodec_free_context.argtypes = [C.POINTER(vp)]
            K.avcodec_get_name.argtypes = [i]
            K.avcodec_get_name.restype = C.c_char_p
            K.av_packet_alloc.restype = vp
            K.av_packet_unref.argtypes = [vp]
            K.av_packet_free.argtypes = [C.POINTER(vp)]
            U.av_frame_alloc.restype = vp
            U.av_frame_unref.argtypes = [vp]
            U.av_frame_free.argtypes = [C.POINTER(vp)]
            U.av_get_pix_fmt_name.argtypes = [i]
            U.av_get_pix_fmt_name.restype = C.c_char_p
            U.av_opt_set_int.argtypes = [vp, C.c_char_p, C.c_int64, i]
            U.av_log_set_level.argtypes = [i]
            U.av_log_set_level(16)                     # AV_LOG_ERROR
            _libs = L
        except Exception:                              # noqa: BLE001  (no cv2, other wheel layout, missing symbol)
            _libs = False
        return _libs


def available() -> bool:
    return bool(_load())


class AvDecoder:
    """Frames of the best video stream of ``path`` in presentation order, one ``next()`` at a time."""

    def __init__(self, path: str, threads: int = 0):
        L = _load()
        if not L:
            raise RuntimeError("libavformat / libavcodec of the cv2 wheel are not usable here")
        self._F, self._K, self._U = L["avformat"], L["avcodec"], L["avutil"]
        self.path = path
        self._fmt, self._ctx, self._pkt, self._frm = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
        self._draining = False
        self.frames_out = 0
        try:
            if self._F.avformat_open_input(C.byref(self._fmt), os.fsencode(path), None, None) < 0:
                raise ValueError(f"{path}: libavformat cannot open this file")
            if self._F.avformat_find_stream_info(self._fmt, None) < 0:
                raise ValueError(f"{path}: no stream information")
            dec = C.c_void_p()
            idx = self._F.av_find_best_stream(self._fmt, _AVMEDIA_TYPE_VIDEO, -1, -1, C.byref(dec), 0)
            if idx < 0 or not dec.value:
                raise ValueError(f"{path}: no decodable video stream")
            # AVFormatContext { av_class, iformat, oformat, priv_data, pb, int ctx_flags, unsigned nb_streams, AVStream **streams }
            nb = C.c_uint.from_address(self._fmt.value + 44).value
            streams = C.c_void_p.from_address(self._fmt.value + 48).value
            if not (0 <= idx < nb <= 256) or not streams:
                raise RuntimeError("AVFormatContext layout check failed")
            st = C.c_void_p.from_address(streams + 8 * idx).value
            # AVStream { av_class, int index, int id, AVCodecParameters *codecpar, ... }
            if not st or C.c_int.from_address(st + 8).value != idx:
                raise RuntimeError("AVStream layout check failed")
            par = C.c_void_p.from_address(st + 16).value
            # AVCodecParameters { enum AVMediaType codec_type, enum AVCodecID codec_id, ... }
            if not par or C.c_int.from_address(par).value != _AVMEDIA_TYPE_VIDEO:
                raise RuntimeError("AVCodecParameters layout check failed")
            self._idx = idx
            self.codec_name = (self._K.avcodec_get_name(C.c_int.from_address(par + 4).value) or b"unknown").decode()
            r = self._F.av_guess_frame_rate(self._fmt, st, None)
            self.fps_num, self.fps_den = (r.num, r.den) if r.num > 0 and r.den > 0 else (30, 1)
            self._ctx = C.c_void_p(self._K.avcodec_alloc_context3(dec))
            if not self._ctx.value or self._K.avcodec_parameters_to_context(self._ctx, par) < 0:
                raise ValueError(f"{path}: cannot set up the decoder")
            # frame threads change nothing in the decoded pictures; 0 lets libavcodec pick from the core count
            self._U.av_opt_set_int(self._ctx, b"threads", threads, 0)
            if self._K.avcodec_open2(self._ctx, dec, None) < 0:
                raise ValueError(f"{path}: cannot open the {self.codec_name} decoder")
            self._pkt = C.c_void_p(self._K.av_packet_alloc())
            self._frm = C.c_void_p(self._U.av_frame_alloc())
            if not self._pkt.value or not self._frm.value:
                raise MemoryError("av_packet_alloc / av_frame_alloc")
            # the first picture tells the geometry and the pixel format (the decoder context's fields are not public ABI)
            self._pending = self._decode()
            if not self._pending:
                raise EOFError(f"{path}: cannot decode the first frame")
            f = self._head()
            self.width, self.height = f.width, f.height
            name = (self._U.av_get_pix_fmt_name(f.format) or b"?").decode()
            if name not in _PIX:
                raise ValueError(f"{path}: unsupported pixel format {name}")
            self.pix_fmt = name
            self.chroma, self.bpc = _PIX[name]
        except Exception:
            self.close()
            raise

    def _head(self) -> _FrameHead:
        return C.cast(self._frm, C.POINTER(_FrameHead)).contents

    def _decode(self) -> bool:
        """One more picture into self._frm; False at the end of the stream."""
        while True:
            rc = self._K.avcodec_receive_frame(self._ctx, self._frm)
            if rc == 0:
                return True
            if rc == _AVERROR_EOF:
                return False
            if rc != _AVERROR_EAGAIN:
                raise IOError(f"{self.path}: decode error {rc}")
            if self._draining:
                return False
            rc = self._F.av_read_frame(self._fmt, self._pkt)
            if rc < 0:                                   # end of file (or a read error): flush the decoder's delay
                self._draining = True
                self._K.avcodec_send_packet(self._ctx, None)
                continue
            if C.cast(self._pkt, C.POINTER(_PacketHead)).contents.stream_index == self._idx:
                self._K.avcodec_send_packet(self._ctx, self._pkt)      # a corrupt packet is skipped, as ffmpeg does
            self._K.av_packet_unref(self._pkt)

    def plane_shapes(self):
        w, h = self.width, self.height
        if self.chroma == 400:
            return [(h, w)]
        cw = (w + 1) // 2 if self.chroma in (420, 422) else w
        ch = (h + 1) // 2 if self.chroma == 420 else h
        return [(h, w), (ch, cw), (ch, cw)]

    def next(self, planes=None, luma_only: bool = False) -> bool:
        """Decodes the next picture; copies it into ``planes`` (arrays shaped like plane_shapes(), u8 or u16) when given.
        False at the end of the stream."""
        if self._pending:
            self._pending = False
        elif not self._decode():
            return False
        if planes is not None:
            f = self._head()
            if (f.width, f.height) != (self.width, self.height) or \
                    (self._U.av_get_pix_fmt_name(f.format) or b"?").decode() != self.pix_fmt:
                raise ValueError(f"{self.path}: picture geometry changes mid-stream (frame {self.frames_out})")
            bps = 1 if self.bpc == 8 else 2
            for k, (ph, pw) in enumerate(self.plane_shapes()):
                if k and luma_only:
                    break
                ls = f.linesize[k]
                if not f.data[k] or ls < pw * bps:
                    raise ValueError(f"{self.path}: unexpected plane layout")
                buf = (C.c_uint8 * (ls * ph)).from_address(f.data[k])
                rows = np.frombuffer(buf, np.uint8).reshape(ph, ls)[:, :pw * bps]
                planes[k][...] = rows if bps == 1 else rows.view("<u2")
        self.frames_out += 1
        return True

    def close(self):
        if getattr(self, "_frm", None) is not None and self._frm.value:
            self._U.av_frame_free(C.byref(self._frm))
        if getattr(self, "_pkt", None) is not None and self._pkt.value:
            self._K.av_packet_free(C.byref(self._pkt))
        if getattr(self, "_ctx", None) is not None and self._ctx.value:
            self._K.avcodec_free_context(C.byref(self._ctx))
        if getattr(self, "_fmt", None) is not None and self._fmt.value:
            self._F.avformat_close_input(C.byref(self._fmt))
        self._frm = self._pkt = self._ctx = self._fmt = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:                              # noqa: BLE001  (interpreter shutdown)
            pass
