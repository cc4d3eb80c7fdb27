"""CPU oracle — TEST INFRASTRUCTURE ONLY.

A restatement of the reference's algorithm for the SuperPoint x2 + LightGlue stereo front-end hot
path, used as the parity checker.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this package; the product path (superslam_b200/) never does and
fails loudly when the CUDA library is missing.

Pinning status:
  * SuperPoint dense network  — PINNED: checked against outputs of the reference's own torch module
    (utils/convert_superpoint_to_onnx.py, weights/superpoint_v1.pth) run in the authoring container;
    fixtures in tests/golden/ made by tests/golden/make_golden.py.
  * FreeList / DescriptorPool handle semantics — PINNED by the reference's own code: oracle/Makefile compiles
    /root/reference/include/DescriptorPool.h + src/DescriptorPool.cc in place into oracle/_ref/libref_pool.so
    (tests/test_oracle_ref_pool.py).
  * stereo post-filter / StereoFrame::backproject / RGB-D post-process — PINNED by the reference's own code:
    src/StereoFrontEnd.cc, src/StereoFrame.cc and src/RgbdFrontEnd.cc compiled in place into
    oracle/_ref/libref_frontend.so (tests/test_oracle_ref_frontend.py; cv::undistortPoints served by cv2).
  * descriptor gather — the reference's real CUDA kernel (src/DescriptorGather.cu) is compiled in place into
    oracle/_ref/libref_gather.so and compared with the product on the GPU (tests/test_gpu_zz_ref_gather.py).
  * keypoint select (border strip, float-vs-double threshold, std::sort tie order, top-K, scale factors) — PINNED by
    the reference's own code: src/SuperPoint.cc is compiled in place (TensorRT reduced to never-called stand-ins,
    oracle/stubs_trt/) into oracle/_ref/libref_nethost.so and its SuperPoint::select_and_gather is run against
    select_keypoints on the reference module's score maps, tie-heavy random maps and threshold-edge values
    (tests/test_oracle_ref_superpoint.py); on the GPU the whole function, pool and gather kernel included, is compared
    with the product's features (tests/test_gpu_zz_ref_gather.py).
  * LightGlue's host half (rows a10, a12: keypoint normalisation, fp32 -> fp16 descriptor binding, matches0 / mscores0
    -> cv::DMatch) — PINNED by the reference's own code: src/LightGlue.cc compiled in place into the same library,
    LightGlue::prepare_inputs / normalize_keypoints / postprocess_outputs run on host buffers
    (tests/test_oracle_ref_lightglue_host.py).
  * the COMPOSITION (image pair -> StereoFrame) — PINNED by the reference's own wrapper code run end to end on the CPU:
    src/SuperPoint.cc, src/LightGlue.cc, src/DescriptorPool.cc and src/StereoFrontEnd.cc compiled in place, unchanged,
    over a functional TensorRT stand-in whose enqueueV3 calls back into the test (the two graphs served by this oracle),
    the CUDA runtime calls on host memory and the gather as a callback (oracle/ref_e2e_shim.cpp ->
    oracle/_ref/libref_e2e.so).  tests/test_oracle_ref_e2e.py: keypoints, responses, fp16 descriptors, match list,
    stereo points and depth flags of the reference run == extract + match + dmatches + stereo_postfilter, bit for bit.
  * the C++ adapter above the C-ABI — EXECUTED under the reference's own caller: src/StereoFrontEnd.cc compiled in
    place over include/superslam_b200_adapter.hpp (oracle/dropin_harness.cpp, functional cv::Mat stand-in in
    oracle/stubs_cv/), against a C-ABI test double on the CPU (oracle/fake_capi.cpp, tests/test_dropin_adapter.py)
    and against the real library on the GPU (tests/test_gpu_zz_dropin.py).
  * LightGlue — PARITY UNPINNED: the arithmetic lives in the un-vendored, un-pinned third-party
    package cvg/LightGlue and no weights are available offline.  The restatement follows the
    published model (lightglue/lightglue.py) and is cross-checked against the independent
    HuggingFace port shipped in this image (transformers 5.5, models/lightglue) with shared weights.
  * EigenPlaces (oracle/eigenplaces.py) — preprocess PINNED bit-exactly against cv2.resize and against the reference's
    own EigenPlaces::preprocess (src/EigenPlaces.cc compiled in place, cv::resize served by cv2; tests/test_oracle_ref_place.py), ResNet18 trunk
    PINNED against torchvision with shared weights, CosineDescriptorIndex / TemporalConsistencyVoter PINNED
    by the reference's own code (src/PlaceRecognizer.cc compiled in place into oracle/_ref/libref_place.so,
    tests/test_oracle_ref_place.py: control flow exact, scores to 1e-6) and by its tests/test_place_recognizer.cc
    (re-expressed); the trained network weights come from
    an un-pinned torch.hub entry and are absent offline: network VALUES are PARITY UNPINNED.
  * image front door (oracle/imgproc.py) - PINNED bit-for-bit against OpenCV: cv2.remap (fixed-point bilinear,
    constant border) and cv2.undistortPoints, live (tests/test_oracle_imgproc.py) and through committed cv2
    outputs (tests/golden/imgproc_cv2.npz, made by tests/golden/make_golden_imgproc.py).
"""
