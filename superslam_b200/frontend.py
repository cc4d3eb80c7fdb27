"""Host-side mirror of the reference's inference interfaces over the C-ABI.

Same names, argument meaning and error behaviour as /root/reference/include/InferenceInterfaces.h:
  SuperPoint.extract / extract_stereo        IFeatureExtractor      (:27-36)
  LightGlue.match / descriptors_to_host      IFeatureMatcher        (:41-59)
  Features / DeviceDescriptors / MatchResult                        (:12-24, DescriptorPool.h:13-20)
  StereoFrontEnd.process                     src/StereoFrontEnd.cc:10-49
The interface methods never raise on a failed inference: like the reference they log and return an
empty Features / MatchResult (src/SuperPoint.cc:895-899, src/LightGlue.cc:381-390).  Constructors do
raise (the reference's initialize() returns false and the facade carries on broken).
Keypoints are numpy arrays instead of cv::KeyPoint: xy float32 [n,2] (pt.x, pt.y) and response [n];
size = 1 and angle = -1 are constants in the reference (src/SuperPoint.cc:716).
"""
from __future__ import annotations

import ctypes as C
import logging
from dataclasses import dataclass, field

import numpy as np

from . import _lib

log = logging.getLogger("superslam_b200")
DESC_DIM = 256


class _Slot:
    """shared_ptr<void> slot_ref (DescriptorPool.h:19,62-76): the last owner returns the slot."""

    def __init__(self, owner: "SuperPoint", slot: int):
        self.owner, self.slot = owner, slot

    def __del__(self):
        try:
            if self.slot >= 0 and self.owner._h:
                _lib.load().ssb_sp_slot_release(self.owner._h, self.slot)
        except Exception:  # interpreter shutdown
            pass


@dataclass
class DeviceDescriptors:
    data: int = 0          # device pointer (fp16 [count, dim] row-major)
    count: int = 0
    dim: int = 0
    slot: int = -1
    slot_ref: object = None
    device: int = 0

    def empty(self) -> bool:
        return self.data == 0 or self.count == 0


@dataclass
class Features:
    keypoints: np.ndarray = field(default_factory=lambda: np.zeros((0, 2), np.float32))
    responses: np.ndarray = field(default_factory=lambda: np.zeros((0,), np.float32))
    descriptors: DeviceDescriptors = field(default_factory=DeviceDescriptors)


@dataclass
class MatchResult:
    """matches: rows (queryIdx, trainIdx); distance = 1 - mscore (src/LightGlue.cc:352-361)."""
    query: np.ndarray = field(default_factory=lambda: np.zeros((0,), np.int32))
    train: np.ndarray = field(default_factory=lambda: np.zeros((0,), np.int32))
    distance: np.ndarray = field(default_factory=lambda: np.zeros((0,), np.float32))
    matches0: np.ndarray = field(default_factory=lambda: np.zeros((0,), np.int32))
    mscores0: np.ndarray = field(default_factory=lambda: np.zeros((0,), np.float32))


def _as_gray_or_bgr(img: np.ndarray):
    img = np.asarray(img)
    if img.dtype != np.uint8:
        raise TypeError("images must be uint8")
    if img.ndim == 2:
        ch = 1
    elif img.ndim == 3 and img.shape[2] in (1, 3):
        ch = img.shape[2]
    else:
        raise ValueError("image must be HxW or HxWx3 (BGR)")
    if img.strides[-1] != 1 or (img.ndim == 3 and img.strides[1] != ch):
        img = np.ascontiguousarray(img)
    return img, ch


class SuperPoint:
    """IFeatureExtractor over ssb_sp_* (reference class SuperPoint, include/SuperPoint.h:37-52)."""

    def __init__(self, weights_path: str, max_keypoints: int, keypoint_threshold: float = 0.005,
                 remove_borders: int = 4, num_slots: int = 8, device: int = 0):
        self._lib = _lib.load()
        self._h = C.c_void_p()
        self.device = device
        self.max_keypoints = max_keypoints
        _lib.check(self._lib.ssb_sp_create(weights_path.encode(), max_keypoints, float(keypoint_threshold),
                                           remove_borders, num_slots, device, C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None):
            self._lib.ssb_sp_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()

    def slots_in_use(self) -> int:
        return self._lib.ssb_sp_slots_in_use(self._h)

    def _extract(self, images):
        prepared = [_as_gray_or_bgr(i) for i in images]
        h, w = prepared[0][0].shape[:2]
        ch = prepared[0][1]
        self._last_ok = False
        for im, c in prepared:
            if im.shape[:2] != (h, w) or c != ch:
                log.error("SuperPoint: stereo pair must share resolution (rectified)")
                return [Features() for _ in images]
        b = len(prepared)
        K = self.max_keypoints
        ptrs = (C.POINTER(C.c_uint8) * b)(*[im.ctypes.data_as(C.POINTER(C.c_uint8)) for im, _ in prepared])
        xy = [np.zeros((K, 2), np.float32) for _ in range(b)]
        sc = [np.zeros((K,), np.float32) for _ in range(b)]
        xyp = (C.POINTER(C.c_float) * b)(*[a.ctypes.data_as(C.POINTER(C.c_float)) for a in xy])
        scp = (C.POINTER(C.c_float) * b)(*[a.ctypes.data_as(C.POINTER(C.c_float)) for a in sc])
        cnt = (C.c_int * b)()
        desc = (C.c_void_p * b)()
        slot = (C.c_int * b)()
        st = self._lib.ssb_sp_extract(self._h, ptrs, b, h, w, prepared[0][0].strides[0], ch, xyp, scp, cnt, desc, slot)
        if st not in (_lib.SSB_OK, _lib.SSB_ERR_EXHAUSTED):
            log.error("SuperPoint: extract failed: %s", self._lib.ssb_last_error().decode())
            return [Features() for _ in images]
        self._last_ok = True
        out = []
        for i in range(b):
            n = cnt[i]
            d = DeviceDescriptors(device=self.device)
            if slot[i] >= 0:
                d = DeviceDescriptors(data=desc[i] or 0, count=n, dim=DESC_DIM, slot=slot[i],
                                      slot_ref=_Slot(self, slot[i]), device=self.device)
            else:
                log.error("SuperPoint: descriptor pool exhausted (no free slot)")
            out.append(Features(keypoints=xy[i][:n].copy(), responses=sc[i][:n].copy(), descriptors=d))
        return out

    def extract(self, image) -> Features:
        return self._extract([image])[0]

    def extract_stereo(self, left, right):
        l, r = self._extract([left, right])
        return l, r

    def infer(self, image):
        """SuperPoint::infer (include/SuperPoint.h:45-46, src/SuperPoint.cc:427-531), the host path of the reference's
        demo programs (tests/test_superpoint_only.cc:71): (ok, keypoints [n,2], responses [n], descriptors f32 [n,256])."""
        f = self.extract(image)
        n = len(f.keypoints)
        desc = np.zeros((n, DESC_DIM), np.float32)
        if n == 0:
            return self._last_ok, f.keypoints, f.responses, desc
        if f.descriptors.empty():
            return False, f.keypoints, f.responses, desc
        st = self._lib.ssb_desc_to_host_f32(self.device, C.c_void_p(f.descriptors.data), n, DESC_DIM,
                                            desc.ctypes.data_as(C.POINTER(C.c_float)))
        return st == _lib.SSB_OK, f.keypoints, f.responses, desc

    def debug_read(self, what: str, shape, dtype):
        out = np.empty(shape, dtype)
        _lib.check(self._lib.ssb_sp_debug_read(self._h, what.encode(), out.ctypes.data_as(C.c_void_p), out.nbytes))
        return out


class LightGlue:
    """IFeatureMatcher over ssb_lg_* (reference class LightGlue, include/LightGlue.h:33-57)."""

    def __init__(self, weights_path: str | None, image_width: int, image_height: int, max_keypoints: int = 1024,
                 device: int = 0, _shared: "LightGlue | None" = None):
        self._lib = _lib.load()
        self._h = C.c_void_p()
        self.device = device
        self.image_width, self.image_height = image_width, image_height
        if _shared is not None:
            _lib.check(self._lib.ssb_lg_clone_context(_shared._h, image_width, image_height, C.byref(self._h)))
            self.max_keypoints = _shared.max_keypoints
        else:
            _lib.check(self._lib.ssb_lg_create(weights_path.encode(), image_width, image_height, max_keypoints,
                                               device, C.byref(self._h)))
            self.max_keypoints = max_keypoints

    @property
    def kp(self) -> int:
        """Rows per image in the workspace (max_keypoints rounded up to 256): the leading dimension of debug buffers."""
        return (self.max_keypoints + 255) // 256 * 256

    def shared_context(self, image_width=None, image_height=None) -> "LightGlue":
        """LightGlue(shared_engine(), w, h): same weights, own stream/workspace (src/SuperSLAM.cc:129-133)."""
        return LightGlue(None, image_width or self.image_width, image_height or self.image_height,
                         device=self.device, _shared=self)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.ssb_lg_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()

    @staticmethod
    def _result(m0, ms0) -> MatchResult:
        q = np.nonzero(m0 >= 0)[0].astype(np.int32)
        return MatchResult(query=q, train=m0[q].astype(np.int32), distance=(np.float32(1.0) - ms0[q]).astype(np.float32),
                           matches0=m0, mscores0=ms0)

    def match(self, kp0, d0, kp1, d1) -> MatchResult:
        """Device path when d0/d1 are DeviceDescriptors, host path when they are float arrays [n,256]."""
        kp0 = np.ascontiguousarray(kp0, np.float32).reshape(-1, 2)
        kp1 = np.ascontiguousarray(kp1, np.float32).reshape(-1, 2)
        n0, n1 = len(kp0), len(kp1)
        m0 = np.full((n0,), -1, np.int32)
        ms0 = np.zeros((n0,), np.float32)
        fp = C.POINTER(C.c_float)
        if isinstance(d0, DeviceDescriptors):
            if d0.empty() or d1.empty() or n0 == 0 or n1 == 0:
                return MatchResult()
            st = self._lib.ssb_lg_match_device(self._h, kp0.ctypes.data_as(fp), n0, C.c_void_p(d0.data),
                                               kp1.ctypes.data_as(fp), n1, C.c_void_p(d1.data),
                                               m0.ctypes.data_as(C.POINTER(C.c_int32)), ms0.ctypes.data_as(fp))
        else:
            if n0 == 0 or n1 == 0:
                return MatchResult()
            a = np.ascontiguousarray(d0, np.float32)
            b = np.ascontiguousarray(d1, np.float32)
            st = self._lib.ssb_lg_match_host(self._h, kp0.ctypes.data_as(fp), n0, a.ctypes.data_as(fp),
                                             kp1.ctypes.data_as(fp), n1, b.ctypes.data_as(fp),
                                             m0.ctypes.data_as(C.POINTER(C.c_int32)), ms0.ctypes.data_as(fp))
        if st != _lib.SSB_OK:
            log.error("LightGlue: match failed: %s", self._lib.ssb_last_error().decode())
            return MatchResult()
        return self._result(m0, ms0)

    def descriptors_to_host(self, d: DeviceDescriptors) -> np.ndarray:
        """CV_32F [count, dim]; empty handle -> empty array (src/LightGlue.cc:460-475)."""
        if d.empty():
            return np.zeros((0, 0), np.float32)
        out = np.empty((d.count, d.dim), np.float32)
        st = self._lib.ssb_desc_to_host_f32(d.device, C.c_void_p(d.data), d.count, d.dim,
                                            out.ctypes.data_as(C.POINTER(C.c_float)))
        if st != _lib.SSB_OK:
            log.error("LightGlue: descriptors_to_host D2H failed")
            return np.zeros((0, 0), np.float32)
        return out

    def debug_read(self, what: str, shape, dtype):
        out = np.empty(shape, dtype)
        _lib.check(self._lib.ssb_lg_debug_read(self._h, what.encode(), out.ctypes.data_as(C.c_void_p), out.nbytes))
        return out


@dataclass
class StereoFrame:
    timestamp: float = 0.0
    keypoints_left: np.ndarray = None
    descriptors_left: DeviceDescriptors = None
    stereo: np.ndarray = None     # [n,3] float64 (uL, uR, v); uR = NaN if no stereo
    has_depth: np.ndarray = None  # [n] int8
    pose_R: np.ndarray = None     # Twc rotation [3,3] (camera-in-world); identity if None
    pose_t: np.ndarray = None     # Twc translation [3]

    def backproject(self, i: int, fx: float, fy: float, cx: float, cy: float, baseline: float) -> np.ndarray:
        """World point of stereo feature i (src/StereoFrame.cc:5-13): Z = fx*b / (uL-uR),
        X = (uL-cx) Z / fx, Y = (v-cy) Z / fy in the camera frame, lifted by the pose Twc."""
        uL, uR, v = (float(x) for x in self.stereo[i])
        Z = fx * baseline / (uL - uR)
        p = np.array([(uL - cx) * Z / fx, (v - cy) * Z / fy, Z], np.float64)
        R = np.eye(3) if self.pose_R is None else np.asarray(self.pose_R, np.float64)
        t = np.zeros(3) if self.pose_t is None else np.asarray(self.pose_t, np.float64)
        return R @ p + t


class StereoFrontEnd:
    """src/StereoFrontEnd.cc:10-49 over any extractor/matcher with the interface methods above."""

    def __init__(self, ext, matcher, min_disparity: float = 1.0):
        self.ext, self.matcher, self.min_disparity = ext, matcher, np.float32(min_disparity)

    def process(self, left, right, timestamp: float = 0.0) -> StereoFrame:
        L, R = self.ext.extract_stereo(left, right)
        n = len(L.keypoints)
        stereo = np.empty((n, 3), np.float64)
        if n:
            stereo[:, 0], stereo[:, 1], stereo[:, 2] = L.keypoints[:, 0], np.nan, L.keypoints[:, 1]
        has_depth = np.zeros((n,), np.int8)
        m = self.matcher.match(L.keypoints, L.descriptors, R.keypoints, R.descriptors)
        for i, j in zip(m.query.tolist(), m.train.tolist()):
            if i < 0 or j < 0 or i >= n or j >= len(R.keypoints):
                continue
            uL, v, uR = L.keypoints[i, 0], L.keypoints[i, 1], R.keypoints[j, 0]
            if np.float32(uL - uR) < self.min_disparity:
                continue
            if abs(np.float32(L.keypoints[i, 1] - R.keypoints[j, 1])) > np.float32(2.0):
                continue
            stereo[i] = (uL, uR, v)
            has_depth[i] = 1
        return StereoFrame(timestamp, L.keypoints, L.descriptors, stereo, has_depth)


class FramePairPipeline:
    """Throughput path (ssb_fe_*): SP x2 + LG + stereo post-filter for `pairs` pairs per call."""

    def __init__(self, sp_weights: str, lg_weights: str, max_keypoints: int, lg_image_width: int,
                 lg_image_height: int, keypoint_threshold: float = 0.005, remove_borders: int = 4,
                 min_disparity: float = 1.0, max_pairs: int = 1, device: int = 0):
        self._lib = _lib.load()
        self._h = C.c_void_p()
        self.K, self.max_pairs, self.device = max_keypoints, max_pairs, device
        _lib.check(self._lib.ssb_fe_create(sp_weights.encode(), lg_weights.encode(), max_keypoints,
                                           float(keypoint_threshold), remove_borders, lg_image_width,
                                           lg_image_height, float(min_disparity), max_pairs, device,
                                           C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None):
            self._lib.ssb_fe_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()

    def _outputs(self, pairs):
        K = self.K
        return dict(count=np.zeros((2 * pairs,), np.int32), xy=np.zeros((2 * pairs, K, 2), np.float32),
                    score=np.zeros((2 * pairs, K), np.float32), matches0=np.zeros((pairs, K), np.int32),
                    mscores0=np.zeros((pairs, K), np.float32), stereo_ur=np.zeros((pairs, K), np.float32),
                    has_depth=np.zeros((pairs, K), np.uint8))

    @staticmethod
    def _ptrs(o):
        fp, ip, i32p, u8p = C.POINTER(C.c_float), C.POINTER(C.c_int), C.POINTER(C.c_int32), C.POINTER(C.c_uint8)
        return (o["count"].ctypes.data_as(ip), o["xy"].ctypes.data_as(fp), o["score"].ctypes.data_as(fp),
                o["matches0"].ctypes.data_as(i32p), o["mscores0"].ctypes.data_as(fp),
                o["stereo_ur"].ctypes.data_as(fp), o["has_depth"].ctypes.data_as(u8p))

    def process(self, images):
        """images: list of 2*pairs gray u8 arrays (left0, right0, left1, ...). One sync per call."""
        images = [np.ascontiguousarray(i, np.uint8) for i in images]
        pairs = len(images) // 2
        h, w = images[0].shape
        ptrs = (C.POINTER(C.c_uint8) * len(images))(*[i.ctypes.data_as(C.POINTER(C.c_uint8)) for i in images])
        o = self._outputs(pairs)
        _lib.check(self._lib.ssb_fe_process(self._h, ptrs, pairs, h, w, w, *self._ptrs(o)))
        return o

    def submit(self, images) -> None:
        """Streaming form of process(): enqueue one step and return; at most two steps in flight.  The
        caller keeps `images` alive until the step is collected (they are copied asynchronously)."""
        images = [np.ascontiguousarray(i, np.uint8) for i in images]
        pairs = len(images) // 2
        h, w = images[0].shape
        ptrs = (C.POINTER(C.c_uint8) * len(images))(*[i.ctypes.data_as(C.POINTER(C.c_uint8)) for i in images])
        _lib.check(self._lib.ssb_fe_submit(self._h, ptrs, pairs, h, w, w))
        self._inflight = getattr(self, "_inflight", [])
        self._inflight.append((images, pairs))

    def collect(self):
        """Results of the oldest submitted step (same dict as process())."""
        images, pairs = self._inflight.pop(0)
        o = self._outputs(pairs)
        n = C.c_int(0)
        _lib.check(self._lib.ssb_fe_collect(self._h, C.byref(n), *self._ptrs(o)))
        assert n.value == pairs
        return o

    def upload(self, images) -> int:
        images = [np.ascontiguousarray(i, np.uint8) for i in images]
        h, w = images[0].shape
        ptrs = (C.POINTER(C.c_uint8) * len(images))(*[i.ctypes.data_as(C.POINTER(C.c_uint8)) for i in images])
        dev = C.c_void_p()
        _lib.check(self._lib.ssb_fe_upload_images(self._h, ptrs, len(images), h, w, w, C.byref(dev)))
        return dev.value

    def enqueue_device(self, images_dev: int, pairs: int, h: int, w: int):
        _lib.check(self._lib.ssb_fe_enqueue_device(self._h, C.c_void_p(images_dev), pairs, h, w))

    def fetch(self, pairs: int):
        o = self._outputs(pairs)
        _lib.check(self._lib.ssb_fe_fetch(self._h, pairs, *self._ptrs(o)))
        return o

    def sync(self):
        _lib.check(self._lib.ssb_fe_sync(self._h))

    def set_rectifiers(self, left: "Rectifier | None", right: "Rectifier | None") -> None:
        """Rectify on the device in front of SuperPoint: calls then take RAW images (image 2p through `left`,
        2p+1 through `right`) - the EuRoC loop of examples/stereo/euroc.cc:176-181 without the host remap."""
        self._rectifiers = (left, right)   # keep them alive: the library borrows the handles
        _lib.check(self._lib.ssb_fe_set_rectifiers(self._h, left._h if left else None, right._h if right else None))

    def set_extract_only(self, on: bool) -> None:
        """SuperPoint only (BASELINE config C1): the 2*pairs images of a call are independent mono frames."""
        _lib.check(self._lib.ssb_fe_set_extract_only(self._h, int(bool(on))))

    # ---- tracking chain (src/VoEstimator.cc:240-246,327) ----
    def enable_tracking(self, on: bool = True) -> None:
        """Every pair slot becomes a stream with a device-resident keyframe; each call also matches that keyframe
        against the current left image (the second LightGlue call of the live pipeline) in the same graph."""
        _lib.check(self._lib.ssb_fe_enable_tracking(self._h, int(bool(on))))

    def reset_tracking(self) -> None:
        _lib.check(self._lib.ssb_fe_reset_tracking(self._h))

    def promote_keyframes(self, pairs: int, mask=None) -> None:
        """`last_keyframe_ = frame` for the streams whose mask entry is set (None: all), on the device."""
        if mask is None:
            _lib.check(self._lib.ssb_fe_promote_keyframes(self._h, None, pairs))
        else:
            m = np.ascontiguousarray(mask, np.uint8)
            assert m.shape == (pairs,)
            _lib.check(self._lib.ssb_fe_promote_keyframes(self._h, m.ctypes.data_as(C.POINTER(C.c_uint8)), pairs))

    def tracking_results(self, pairs: int):
        """Tracking outputs of the step delivered last: keyframe_count [pairs], track_matches0 / track_mscores0 /
        track_usable [pairs, K] (queryIdx = keyframe feature, trainIdx = current left feature)."""
        K = self.K
        o = dict(keyframe_count=np.zeros((pairs,), np.int32), track_matches0=np.zeros((pairs, K), np.int32),
                 track_mscores0=np.zeros((pairs, K), np.float32), track_usable=np.zeros((pairs, K), np.uint8))
        _lib.check(self._lib.ssb_fe_tracking_results(
            self._h, pairs, o["keyframe_count"].ctypes.data_as(C.POINTER(C.c_int)),
            o["track_matches0"].ctypes.data_as(C.POINTER(C.c_int32)),
            o["track_mscores0"].ctypes.data_as(C.POINTER(C.c_float)),
            o["track_usable"].ctypes.data_as(C.POINTER(C.c_uint8))))
        return o

    def kernel_launches_per_call(self, pairs: int) -> int:
        return int(self._lib.ssb_fe_kernel_launches_per_call(self._h, pairs))

    def event_record(self, idx: int):
        _lib.check(self._lib.ssb_fe_event_record(self._h, idx))

    def event_elapsed_ms(self, a: int, b: int) -> float:
        ms = C.c_float()
        _lib.check(self._lib.ssb_fe_event_elapsed_ms(self._h, a, b, C.byref(ms)))
        return ms.value

    @property
    def kp(self) -> int:
        """Rows per image in the LightGlue workspace (K rounded up to 256): leading dimension of its debug buffers."""
        return (self.K + 255) // 256 * 256

    def lightglue_debug_read(self, what, shape, dtype):
        out = np.empty(shape, dtype)
        h = self._lib.ssb_fe_lightglue(self._h)
        _lib.check(self._lib.ssb_lg_debug_read(C.c_void_p(h), what.encode(), out.ctypes.data_as(C.c_void_p), out.nbytes))
        return out

    def superpoint_debug_read(self, what, shape, dtype):
        out = np.empty(shape, dtype)
        h = self._lib.ssb_fe_superpoint(self._h)
        _lib.check(self._lib.ssb_sp_debug_read(C.c_void_p(h), what.encode(), out.ctypes.data_as(C.c_void_p), out.nbytes))
        return out


class MultiGpuFrontEnd:
    """ssb_mg_*: the frame-pair front end on several GPUs of one process (one host thread + one front end per device,
    pair p on device p mod G, no data-path collective).  Same outputs as FramePairPipeline.process, any number of pairs."""

    def __init__(self, sp_weights: str, lg_weights: str, max_keypoints: int, lg_image_width: int, lg_image_height: int,
                 devices, max_pairs_per_device: int = 8, keypoint_threshold: float = 0.005, remove_borders: int = 4,
                 min_disparity: float = 1.0):
        self._lib = _lib.load()
        self._h = C.c_void_p()
        self.K = max_keypoints
        ids = (C.c_int * len(devices))(*devices)
        _lib.check(self._lib.ssb_mg_create(sp_weights.encode(), lg_weights.encode(), max_keypoints,
                                           float(keypoint_threshold), remove_borders, lg_image_width, lg_image_height,
                                           float(min_disparity), max_pairs_per_device, ids, len(devices),
                                           C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None):
            self._lib.ssb_mg_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()

    def device_of_pair(self, pair: int) -> int:
        return int(self._lib.ssb_mg_device_of_pair(self._h, pair))

    def process(self, images):
        images = [np.ascontiguousarray(i, np.uint8) for i in images]
        pairs = len(images) // 2
        h, w = images[0].shape
        ptrs = (C.POINTER(C.c_uint8) * len(images))(*[i.ctypes.data_as(C.POINTER(C.c_uint8)) for i in images])
        K = self.K
        o = dict(count=np.zeros((2 * pairs,), np.int32), xy=np.zeros((2 * pairs, K, 2), np.float32),
                 score=np.zeros((2 * pairs, K), np.float32), matches0=np.zeros((pairs, K), np.int32),
                 mscores0=np.zeros((pairs, K), np.float32), stereo_ur=np.zeros((pairs, K), np.float32),
                 has_depth=np.zeros((pairs, K), np.uint8))
        st = self._lib.ssb_mg_process(self._h, ptrs, pairs, h, w, w, *FramePairPipeline._ptrs(o))
        if st != _lib.SSB_OK:
            raise _lib.SsbError(st, (self._lib.ssb_mg_last_error(self._h) or b"").decode())
        return o


def kernel_launch_count() -> int:
    return int(_lib.load().ssb_kernel_launch_count())


class EigenPlaces:
    """IPlaceRecognizer over ssb_ep_* (reference class EigenPlaces, include/EigenPlaces.h:21-66;
    interface include/PlaceRecognizer.h:20-36).  compute_global_descriptor never raises on a failed
    inference: like the reference (src/EigenPlaces.cc:146-147,157-160) it logs and returns an empty row."""

    def __init__(self, weights_path: str, input_width: int, input_height: int, max_batch: int = 1,
                 device: int = 0, min_score: float | None = None):
        import os

        self._lib = _lib.load()
        self._h = C.c_void_p()
        self.input_width, self.input_height = input_width, input_height
        # EigenPlaces::EigenPlaces reads SUPERSLAM_LOOP_MIN_SCORE (src/EigenPlaces.cc:30-34); default 0.75
        env = os.environ.get("SUPERSLAM_LOOP_MIN_SCORE")
        self.min_score = float(min_score) if min_score is not None else (float(env) if env else 0.75)
        _lib.check(self._lib.ssb_ep_create(weights_path.encode(), input_width, input_height, max_batch, device,
                                           C.byref(self._h)))
        self.dim = self._lib.ssb_ep_descriptor_dim(self._h)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.ssb_ep_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()

    def compute_global_descriptors(self, images) -> np.ndarray:
        """Batch form: same-size images -> [n, 512] float32 (rows L2-normalised); empty on failure."""
        prepared = [_as_gray_or_bgr(i) for i in images]
        if not prepared:
            return np.zeros((0, self.dim), np.float32)
        h, w = prepared[0][0].shape[:2]
        ch = prepared[0][1]
        if any(im.shape[:2] != (h, w) or c != ch for im, c in prepared):
            log.error("EigenPlaces: images of one call must share size and channel count")
            return np.zeros((0, self.dim), np.float32)
        n = len(prepared)
        ptrs = (C.POINTER(C.c_uint8) * n)(*[im.ctypes.data_as(C.POINTER(C.c_uint8)) for im, _ in prepared])
        out = np.zeros((n, self.dim), np.float32)
        st = self._lib.ssb_ep_compute(self._h, ptrs, n, h, w, prepared[0][0].strides[0], ch,
                                      out.ctypes.data_as(C.POINTER(C.c_float)))
        if st != _lib.SSB_OK:
            log.error("EigenPlaces: inference failed: %s", self._lib.ssb_last_error().decode())
            return np.zeros((0, self.dim), np.float32)
        return out

    def compute_global_descriptor(self, image) -> np.ndarray:
        """[1, 512] float32, L2-normalised; empty [0, 512] on failure."""
        return self.compute_global_descriptors([image])

    def add(self, keyframe_id: int, global_descriptor) -> None:
        d = np.ascontiguousarray(global_descriptor, np.float32).reshape(-1)
        _lib.check(self._lib.ssb_ep_add(self._h, int(keyframe_id), d.ctypes.data_as(C.POINTER(C.c_float)), len(d)))

    def query(self, global_descriptor, exclude_recent: int, top_k: int, min_score: float | None = None):
        """-> list of (keyframe_id, score) by descending score (LoopCandidate, PlaceRecognizer.h:11-16)."""
        d = np.ascontiguousarray(global_descriptor, np.float32).reshape(-1)
        cap = max(1, self.size() if top_k <= 0 else min(top_k, max(1, self.size())))
        ids = (C.c_uint64 * cap)()
        sc = np.zeros((cap,), np.float32)
        n = C.c_int(0)
        ms = self.min_score if min_score is None else float(min_score)
        _lib.check(self._lib.ssb_ep_query(self._h, d.ctypes.data_as(C.POINTER(C.c_float)), len(d), int(exclude_recent),
                                          int(top_k), C.c_float(ms), ids, sc.ctypes.data_as(C.POINTER(C.c_float)), cap,
                                          C.byref(n)))
        return [(int(ids[i]), float(sc[i])) for i in range(n.value)]

    def size(self) -> int:
        return self._lib.ssb_ep_index_size(self._h)

    def debug_read(self, what: str, shape, dtype):
        out = np.empty(shape, dtype)
        _lib.check(self._lib.ssb_ep_debug_read(self._h, what.encode(), out.ctypes.data_as(C.c_void_p), out.nbytes))
        return out


class Rectifier:
    """cv::remap(img, out, M1, M2, cv::INTER_LINEAR) on the device (ssb_rect_*): the stereo rectification the
    reference's EuRoC driver applies before track_stereo (examples/stereo/euroc.cc:118-133,176-177).
    `map_x`, `map_y` are the CV_32F maps of cv::initUndistortRectifyMap for one camera."""

    def __init__(self, map_x, map_y, src_shape, max_images: int = 2, device: int = 0):
        self._lib = _lib.load()
        self._h = C.c_void_p()
        mx = np.ascontiguousarray(map_x, np.float32)
        my = np.ascontiguousarray(map_y, np.float32)
        if mx.shape != my.shape or mx.ndim != 2:
            raise ValueError("map_x / map_y must be 2-D arrays of one shape")
        self.dst_shape = mx.shape
        self.src_shape = tuple(src_shape[:2])
        _lib.check(self._lib.ssb_rect_create(mx.ctypes.data_as(C.POINTER(C.c_float)),
                                             my.ctypes.data_as(C.POINTER(C.c_float)), mx.shape[0], mx.shape[1],
                                             self.src_shape[0], self.src_shape[1], max_images, device,
                                             C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None):
            self._lib.ssb_rect_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()

    def remap(self, images):
        """list of gray u8 images (src_shape) -> list of rectified u8 images (dst_shape)."""
        imgs = [np.asarray(i) for i in images]
        for im in imgs:
            if im.dtype != np.uint8 or im.ndim != 2 or im.shape != self.src_shape or im.strides[1] != 1:
                raise ValueError("remap expects gray u8 images of the source shape")
        n = len(imgs)
        outs = [np.empty(self.dst_shape, np.uint8) for _ in range(n)]
        ip = (C.POINTER(C.c_uint8) * n)(*[im.ctypes.data_as(C.POINTER(C.c_uint8)) for im in imgs])
        op = (C.POINTER(C.c_uint8) * n)(*[o.ctypes.data_as(C.POINTER(C.c_uint8)) for o in outs])
        _lib.check(self._lib.ssb_rect_remap(self._h, ip, n, imgs[0].strides[0], op))
        return outs

    def remap_device(self, src_dev: int, count: int, dst_dev: int) -> None:
        _lib.check(self._lib.ssb_rect_remap_device(self._h, C.c_void_p(src_dev), count, C.c_void_p(dst_dev)))


class RgbdFrontEnd:
    """include/RgbdFrontEnd.h:15-37 / src/RgbdFrontEnd.cc:24-58 over any extractor with `extract`:
    keypoint undistortion, depth sampling and the synthetic right coordinate run on the device
    (ssb_rgbd_process)."""

    def __init__(self, ext, fx: float, fy: float, cx: float, cy: float, baseline: float, depth_factor: float,
                 max_depth: float, dist_coeffs=None, max_keypoints: int = 4096, max_shape=(1080, 1920),
                 device: int = 0):
        self._lib = _lib.load()
        self.ext = ext
        self.camera = (C.c_double * 4)(fx, fy, cx, cy)
        d = np.zeros((0,), np.float64) if dist_coeffs is None else np.asarray(dist_coeffs, np.float64).ravel()
        self.dist = d
        self.bf = float(fx) * float(baseline)   # K_.fx() * K_.baseline()
        self.depth_factor, self.max_depth = float(depth_factor), float(max_depth)
        self._h = C.c_void_p()
        _lib.check(self._lib.ssb_rgbd_create(max_keypoints, max_shape[0], max_shape[1], device, C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None):
            self._lib.ssb_rgbd_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()

    def postprocess(self, keypoints, depth):
        """raw keypoints [n,2] + depth image (uint16 or float32) -> (undistorted xy, stereo [n,3], has_depth)."""
        xy = np.ascontiguousarray(keypoints, np.float32).reshape(-1, 2)
        n = len(xy)
        depth = np.asarray(depth)
        if depth.dtype == np.uint16:
            dtype = 0
        elif depth.dtype == np.float32:
            dtype = 1
        else:   # sampleDepth returns 0 for any other type: no depth anywhere
            depth, dtype = np.zeros(depth.shape[:2], np.uint16), 0
        if depth.strides[1] != depth.itemsize:
            depth = np.ascontiguousarray(depth)
        oxy = np.empty((n, 2), np.float32)
        stereo = np.empty((n, 3), np.float64)
        has = np.zeros((n,), np.uint8)
        dp = self.dist.ctypes.data_as(C.POINTER(C.c_double)) if len(self.dist) else None
        _lib.check(self._lib.ssb_rgbd_process(
            self._h, xy.ctypes.data_as(C.POINTER(C.c_float)), n, depth.ctypes.data_as(C.c_void_p), dtype,
            depth.shape[0], depth.shape[1], depth.strides[0], self.camera, dp, len(self.dist), self.bf,
            self.depth_factor, self.max_depth, oxy.ctypes.data_as(C.POINTER(C.c_float)),
            stereo.ctypes.data_as(C.POINTER(C.c_double)), has.ctypes.data_as(C.POINTER(C.c_uint8))))
        return oxy, stereo, has.astype(np.int8)

    def process(self, gray, depth, timestamp: float = 0.0) -> StereoFrame:
        L = self.ext.extract(gray)
        oxy, stereo, has = self.postprocess(L.keypoints, depth)
        return StereoFrame(timestamp, oxy, L.descriptors, stereo, has)
