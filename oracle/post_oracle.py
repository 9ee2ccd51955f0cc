"""O-cpu for the frame post passes.  TEST INFRASTRUCTURE: imported by tests/ only.

numpy restatement of tsd/src/render_pipeline/passes:
  convert_float_color   convertFloatColorBuffer      AnariSceneRenderPass.cpp:15-20
  composite_depth       compositeFrame               AnariSceneRenderPass.cpp:30-46
  outline               computeOutline + shadePixel  OutlineRenderPass.cpp:13-46
  visualize_depth       computeDepthImage            VisualizeDepthPass.cpp:13-21
PINNED (tests/test_post_ref_host.py): every function here is checked bit for bit against the reference's own pass
sources compiled in place into oracle/_ref/libref_post.so (oracle/ref_post/, recipe in oracle/Makefile) — loop bounds,
the unsigned window arithmetic of computeOutline, comparison / clamp / truncation semantics are reference code.
Still restated, in the shim that build uses as well as here: helium::cvt_color_to_float4 / cvt_color_to_uint32 and
linalg's lerp live in the ANARI-SDK (helium/helium_math.h >= 0.15, anari_cpp/ext/linalg.h), which is not vendored with
the reference: c/255.f per byte, uint32(255.f * clamp(f,0,1)) per component, r|g<<8|b<<16|a<<24; lerp(a,b,t) =
a*(1-t) + b*t, every operation rounded to fp32 (parity unpinned for these three helpers only).
"""
import numpy as np

F = np.float32


def _cvt_component(f):
    return (F(255.0) * np.clip(f.astype(F), F(0), F(1))).astype(F).astype(np.uint32)


def convert_float_color(rgba_f32):
    v = (np.clip(np.asarray(rgba_f32, F).reshape(-1, 4), F(0), F(1)) * F(255.0)).astype(F).astype(np.uint8).astype(np.uint32)
    return v[:, 0] | (v[:, 1] << 8) | (v[:, 2] << 16) | (v[:, 3] << 24)


def composite_depth(color_out, depth_out, id_out, color_in, depth_in, id_in, first_pass):
    take = np.ones_like(depth_in, bool) if first_pass else depth_in < depth_out
    color_out, depth_out = color_out.copy(), depth_out.copy()
    color_out[take] = color_in[take]
    depth_out[take] = depth_in[take]
    if id_in is not None:
        id_out = id_out.copy()
        id_out[take] = id_in[take]
    return color_out, depth_out, id_out


def shade_pixel(c):
    c = np.asarray(c, np.uint32)
    hl = (F(1.0), F(0.5), F(0.0), F(1.0))
    out = np.zeros_like(c)
    for k in range(4):
        cin = (((c >> (8 * k)) & 0xFF).astype(F) / F(255.0)).astype(F)
        v = ((cin * (F(1.0) - F(0.8))).astype(F) + F(hl[k] * F(0.8))).astype(F)
        out |= _cvt_component(v) << np.uint32(8 * k)
    return out


def outline(color, object_id, width, height, outline_id):
    """The reference computes the window start as max(0u, y - 1) in UNSIGNED arithmetic: on row 0 (column 0) that
    wraps to UINT_MAX and the window is empty, so those pixels are never outlined."""
    color = np.asarray(color, np.uint32).reshape(height, width).copy()
    ids = np.asarray(object_id, np.uint32).reshape(height, width) == np.uint32(outline_id)
    cnt = np.zeros((height, width), np.int32)
    for y in range(1, height):
        y1 = min(height - 1, y + 1)
        for x in range(1, width):
            x1 = min(width - 1, x + 1)
            cnt[y, x] = ids[y - 1:y1 + 1, x - 1:x1 + 1].sum()
    sel = (cnt > 1) & (cnt < 8)
    color[sel] = shade_pixel(color[sel])
    return color.ravel()


def visualize_depth(depth, max_depth):
    v = np.clip((np.asarray(depth, F) / F(max_depth)).astype(F), F(0), F(1))
    c = _cvt_component(v)
    return c | (c << 8) | (c << 16) | np.uint32(255 << 24)
