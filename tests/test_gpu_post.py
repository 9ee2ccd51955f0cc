"""GPU suite: the frame post passes of include/dvr_b200.h (dvr_post_*) against the numpy restatement of TSD's render
pipeline passes (oracle/post_oracle.py) — bit-exact (byte / integer work), on random buffers and on the channels of a
real frame mapped through ANARI_NV_FRAME_BUFFERS_CUDA."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import post_oracle as PO  # noqa: E402
from test_gpu_anari import AnariScene  # noqa: E402
from visrtx_b200 import anari as A  # noqa: E402
from visrtx_b200 import capi  # noqa: E402

pytestmark = pytest.mark.gpu


def _dev(a):
    import torch
    t = torch.from_numpy(np.ascontiguousarray(a).view(np.int32) if a.dtype == np.uint32 else np.ascontiguousarray(a))
    return t.cuda()


def _host(t, dtype=None):
    a = t.cpu().numpy()
    return a.view(np.uint32) if dtype == np.uint32 else a


@pytest.mark.parametrize("n", [1, 255, 1920 * 1080])
def test_convert_float_color(n):
    rng = np.random.default_rng(n)
    src = (rng.random((n, 4)) * 1.4 - 0.2).astype(np.float32)
    src[0] = (0.0, 1.0, 0.999999, 2.0)
    d_in, d_out = _dev(src), _dev(np.zeros(n, np.uint32))
    capi.post_convert_float_color(d_in.data_ptr(), d_out.data_ptr(), n)
    assert np.array_equal(_host(d_out, np.uint32), PO.convert_float_color(src))


@pytest.mark.parametrize("first,with_ids", [(True, True), (False, True), (False, False)])
def test_composite_depth(first, with_ids):
    rng = np.random.default_rng(5)
    n = 640 * 480 + 3
    co, ci = rng.integers(0, 2**32, n, dtype=np.uint32), rng.integers(0, 2**32, n, dtype=np.uint32)
    do, di = rng.random(n).astype(np.float32), rng.random(n).astype(np.float32)
    do[::7] = np.inf
    io, ii = rng.integers(0, 50, n, dtype=np.uint32), rng.integers(0, 50, n, dtype=np.uint32)
    dco, ddo, dio, dci, ddi, dii = (_dev(x) for x in (co, do, io, ci, di, ii))
    capi.post_composite_depth(dco.data_ptr(), ddo.data_ptr(), dio.data_ptr() if with_ids else 0, dci.data_ptr(),
                              ddi.data_ptr(), dii.data_ptr() if with_ids else 0, n, first)
    wc, wd, wi = PO.composite_depth(co, do, io, ci, di, ii if with_ids else None, first)
    assert np.array_equal(_host(dco, np.uint32), wc) and np.array_equal(_host(ddo), wd)
    assert np.array_equal(_host(dio, np.uint32), wi)


@pytest.mark.parametrize("w,h", [(64, 48), (1, 1), (3, 2), (257, 131)])
def test_outline(w, h):
    rng = np.random.default_rng(w * h)
    ids = np.full((h, w), 0xFFFFFFFF, np.uint32)
    yy, xx = np.mgrid[0:h, 0:w]
    ids[(yy - h / 2) ** 2 + (xx - w / 3) ** 2 < (min(w, h) / 3) ** 2] = 7  # a disc touching row/column 0 when small
    ids[: max(h // 5, 1), : max(w // 4, 1)] = 7  # a block in the corner: row 0 and column 0 are never outlined
    color = rng.integers(0, 2**32, w * h, dtype=np.uint32)
    dc, di = _dev(color), _dev(ids.ravel())
    capi.post_outline(dc.data_ptr(), di.data_ptr(), w, h, 7)
    got, want = _host(dc, np.uint32), PO.outline(color, ids.ravel(), w, h, 7)
    assert np.array_equal(got, want)
    if w > 8:
        changed = (got != color).reshape(h, w)
        assert changed.any() and not changed[0].any() and not changed[:, 0].any()
    # outlineId ~0u disables the pass (OutlineRenderPass.cpp:58)
    dc2 = _dev(color)
    capi.post_outline(dc2.data_ptr(), di.data_ptr(), w, h, 0xFFFFFFFF)
    assert np.array_equal(_host(dc2, np.uint32), color)


def test_visualize_depth_and_pick_on_a_real_frame():
    import torch
    s = AnariScene(40, 96, 64, "raycast", 0.5, channels=("depth", "objectId"))
    s.render()
    host_depth, w, h, _ = s.d.map_frame(s.frame, "channel.depth")
    host_ids, _, _, _ = s.d.map_frame(s.frame, "channel.objectId")
    dptr, _, _, _ = s.d.map_frame(s.frame, "channel.depthCUDA")
    iptr, _, _, _ = s.d.map_frame(s.frame, "channel.objectIdCUDA")
    cptr, _, _, _ = s.d.map_frame(s.frame, "channel.colorCUDA")
    out = torch.zeros(w * h, dtype=torch.int32, device="cuda")
    capi.post_visualize_depth(out.data_ptr(), dptr, w * h, 6.0)
    want = PO.visualize_depth(np.asarray(host_depth).ravel(), 6.0)
    assert np.array_equal(_host(out, np.uint32), want)
    assert len(np.unique(want)) > 10  # a real depth ramp, not a constant
    # pick: the volume (id 7) in the centre, nothing in the corner
    d_c, id_c = capi.post_pick(dptr, iptr, w, h, w // 2, h // 2)
    assert id_c == 7 and d_c == np.asarray(host_depth).reshape(h, w)[h // 2, w // 2]
    d_0, id_0 = capi.post_pick(dptr, iptr, w, h, 0, 0)
    assert id_0 == np.asarray(host_ids).reshape(h, w)[0, 0] and d_0 == np.asarray(host_depth).reshape(h, w)[0, 0]
    # outline of the picked object on the frame's own colour buffer == oracle on the host copies
    host_color, _, _, _ = s.d.map_frame(s.frame, "channel.color")
    before = np.array(host_color, copy=True).ravel()
    capi.post_outline(cptr, iptr, w, h, 7)
    torch.cuda.synchronize()
    after = torch.empty(w * h, dtype=torch.int32, device="cuda")
    import ctypes as C
    C.CDLL("libcudart.so").cudaMemcpy(C.c_void_p(after.data_ptr()), C.c_void_p(cptr), C.c_size_t(w * h * 4), C.c_int(3))
    want = PO.outline(before, np.asarray(host_ids).ravel(), w, h, 7)
    assert np.array_equal(_host(after, np.uint32), want) and (want != before).sum() > 20
    s.close()


def test_tsd_render_pipeline_flow_on_the_device():
    """The call sequence of TSD's AnariSceneRenderPass (tsd/src/render_pipeline/passes/AnariSceneRenderPass.cpp:
    79-96 constructor, 139-166 setEnableIDs, 207-262 render / copyFrameData) followed by the outline pass, on the
    CUDA channel maps: `accumulation` frame parameter, channel.objectId switched on and off with
    anariUnsetParameter, anariDiscardFrame + wait, anariFrameReady(NO_WAIT) polling, depth-tested composite."""
    import ctypes as C
    import torch
    s = AnariScene(40, 96, 64, "default", 0.5, channels=("depth",))
    d, f = s.d, s.frame
    d.set(f, "accumulation", A.BOOL, 1)  # AnariSceneRenderPass.cpp:85 (accepted; accumulation is always on)
    d.commit(f)
    w, h = 96, 64
    n = w * h
    inf = float("inf")
    pipe_color = torch.zeros(n, dtype=torch.int32, device="cuda")
    pipe_depth = torch.full((n,), inf, dtype=torch.float32, device="cuda")
    pipe_ids = torch.full((n,), -1, dtype=torch.int32, device="cuda")

    def copy_frame_data(enable_ids):
        cptr, cw, ch, ct = d.map_frame(f, "channel.colorCUDA")
        dptr, _, _, _ = d.map_frame(f, "channel.depthCUDA")
        assert (cw, ch, ct) == (w, h, A.UFIXED8_RGBA_SRGB)
        iptr = d.map_frame(f, "channel.objectIdCUDA")[0] if enable_ids else None
        capi.post_composite_depth(pipe_color.data_ptr(), pipe_depth.data_ptr(), pipe_ids.data_ptr() if iptr else 0,
                                  cptr, dptr, iptr or 0, n, True)
        torch.cuda.synchronize()
        return iptr

    # first frame: render + wait, then poll-and-resubmit like render() does
    d.render(f)
    d.wait(f)
    assert A.lib.anariFrameReady(d.handle, f, A.NO_WAIT) == 1
    assert copy_frame_data(False) is None
    d.render(f)
    # enable IDs: discard + wait, set the channel, commit, render + wait (setEnableIDs(true))
    A.lib.anariDiscardFrame(d.handle, f)
    d.wait(f)
    d.set(f, "channel.objectId", A.DATA_TYPE, A.UINT32)
    d.commit(f)
    d.render(f)
    d.wait(f)
    assert copy_frame_data(True)
    ids = _host(pipe_ids, np.uint32)
    assert set(np.unique(ids)) == {7, 0xFFFFFFFF}
    before = _host(pipe_color, np.uint32).copy()
    capi.post_outline(pipe_color.data_ptr(), pipe_ids.data_ptr(), w, h, 7)
    torch.cuda.synchronize()
    assert np.array_equal(_host(pipe_color, np.uint32), PO.outline(before, ids, w, h, 7))
    # disable IDs again: the channel disappears from the frame (setEnableIDs(false))
    d.unset(f, "channel.objectId")
    d.commit(f)
    d.render(f)
    d.wait(f)
    ptr, _, _, t = d.map_frame(f, "channel.objectIdCUDA")
    assert ptr is None and t == A.UNKNOWN
    assert not [m for m in d.messages if m[0] <= A.SEVERITY_ERROR], d.messages
    s.close()
