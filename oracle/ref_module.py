"""TEST INFRASTRUCTURE ONLY -- torch-facing wrapper of oracle/_ref/libadgs_ref.so, i.e. the
UNMODIFIED reference rasterizer + simple-knn compiled from /root/reference by oracle/Makefile.

Re-creates what the reference's torch binding does around CudaRasterizer::Rasterizer
(submodules/depth-diff-gaussian-rasterization/rasterize_points.cu:35-275: output allocation,
zero-fill, resizable byte arenas) so the library can be driven with the same tensors as
adgs_b200.rasterizer._C. Needs a GPU. Never imported by the product path.
"""
import ctypes as C
import os
import subprocess

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libadgs_ref.so")
REFERENCE_ROOT = "/root/reference"

_ALLOC = C.CFUNCTYPE(C.c_void_p, C.c_size_t, C.c_void_p)
_lib = None


def available() -> bool:
    return os.path.exists(LIB_PATH)


def build() -> bool:
    """Compile the reference from /root/reference (only possible where that checkout exists)."""
    if not os.path.isdir(REFERENCE_ROOT):
        return available()
    res = subprocess.run(["make", "-j8", "-C", _HERE], capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("building oracle/_ref failed:\n" + res.stdout + res.stderr)
    return True


class GeomLayout(C.Structure):
    _fields_ = [(n, C.c_size_t) for n in ("depths", "clamped", "internal_radii", "means2D", "cov3D", "conic_opacity",
                                          "rgb", "tiles_touched", "point_offsets", "total")]


class BinningLayout(C.Structure):
    _fields_ = [(n, C.c_size_t) for n in ("point_list", "point_list_unsorted", "point_list_keys",
                                          "point_list_keys_unsorted", "total")]


class ImageLayout(C.Structure):
    _fields_ = [(n, C.c_size_t) for n in ("n_contrib", "ranges", "total")]


def load():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError(f"{LIB_PATH} missing: run `make -C oracle` where /root/reference exists")
        _lib = C.CDLL(LIB_PATH)
        _lib.ref_forward.restype = C.c_int
        _lib.ref_backward.restype = C.c_int
    return _lib


def _join_default_stream(dev):
    """The reference launches on the legacy default stream (stream 0), which is also torch's default stream: kernels
    queued by torch before / after the call are ordered with it without any host synchronisation -- exactly like
    the stock binding, whose only host block is the 4-byte num_rendered read (rasterizer_impl.cu:288). Only a caller
    running torch on a NON-default stream needs an explicit join."""
    if torch.cuda.current_stream(dev).cuda_stream != 0:
        torch.cuda.synchronize(dev)


def _p(t):
    if t is None or t.numel() == 0:
        return C.c_void_p(None)
    return C.c_void_p(t.data_ptr())


def _c(t):
    return t if (t is None or t.numel() == 0) else t.contiguous()


def rasterize_gaussians(background, means3D, colors, opacity, scales, rotations, scale_modifier, cov3D_precomp,
                        viewmatrix, projmatrix, tan_fovx, tan_fovy, image_height, image_width, sh, flow_points,
                        semantic, degree, campos, prefiltered, inv_depth, debug):
    """Same signature / result tuple as the reference's _C.rasterize_gaussians."""
    lib = load()
    dev = means3D.device
    P, H, W = means3D.shape[0], int(image_height), int(image_width)
    D_S = semantic.shape[1] if (semantic is not None and semantic.numel()) else 0
    M = sh.shape[1] if (sh is not None and sh.numel()) else 0
    o = dict(dtype=torch.float32, device=dev)
    out_color = torch.zeros((3, H, W), **o)
    out_depth = torch.zeros((1, H, W), **o)
    img_opacity = torch.zeros((1, H, W), **o)
    img_flow = torch.zeros((3, H, W), **o)
    img_semantic = torch.zeros((D_S, H, W), **o)
    radii = torch.zeros((P,), dtype=torch.int32, device=dev)
    bufs = {}

    def mk(name):
        def f(n, _u):
            bufs[name] = torch.empty((int(n),), dtype=torch.uint8, device=dev)
            return bufs[name].data_ptr()
        return _ALLOC(f)

    cbs = [mk("geom"), mk("binning"), mk("img")]
    keep = [_c(t) for t in (background, means3D, sh, colors, flow_points, semantic, opacity, scales, rotations,
                            cov3D_precomp, viewmatrix, projmatrix, campos)]
    (background, means3D, sh, colors, flow_points, semantic, opacity, scales, rotations, cov3D_precomp, viewmatrix,
     projmatrix, campos) = keep
    rendered = 0
    if P != 0:
        _join_default_stream(dev)
        with torch.cuda.device(dev):
            rendered = lib.ref_forward(
                cbs[0], cbs[1], cbs[2], None, C.c_int(P), C.c_int(int(degree)), C.c_int(M), C.c_int(D_S),
                _p(background), C.c_int(W), C.c_int(H), _p(means3D), _p(sh), _p(colors), _p(flow_points),
                _p(semantic), _p(opacity), _p(scales), C.c_float(scale_modifier), _p(rotations), _p(cov3D_precomp),
                _p(viewmatrix), _p(projmatrix), _p(campos), C.c_float(tan_fovx), C.c_float(tan_fovy),
                C.c_int(int(prefiltered)), _p(out_color), _p(out_depth), _p(img_opacity), _p(img_flow),
                _p(img_semantic), C.c_int(int(inv_depth)), _p(radii), C.c_int(int(debug)))
        _join_default_stream(dev)
        if rendered < 0:
            raise RuntimeError("reference forward failed")
    e = torch.empty((0,), dtype=torch.uint8, device=dev)
    return (rendered, out_color, out_depth, img_opacity, radii, bufs.get("geom", e), bufs.get("binning", e),
            bufs.get("img", e), img_flow, img_semantic)


def rasterize_gaussians_backward(background, means3D, radii, colors, scales, rotations, scale_modifier,
                                 cov3D_precomp, viewmatrix, projmatrix, tan_fovx, tan_fovy, dL_dout_color,
                                 dL_dout_depth, dL_dout_flow, dL_dout_semantic, semantic, flow_points, sh, degree,
                                 campos, geomBuffer, R, binningBuffer, imageBuffer, img_opacity, grad_img_opacity,
                                 inv_depth, debug):
    """Same signature / result tuple as the reference's _C.rasterize_gaussians_backward."""
    lib = load()
    dev = means3D.device
    P = means3D.shape[0]
    H, W = dL_dout_color.shape[1], dL_dout_color.shape[2]
    M = sh.shape[1] if (sh is not None and sh.numel()) else 0
    D_S = semantic.shape[1] if (semantic is not None and semantic.numel()) else 0
    o = dict(dtype=torch.float32, device=dev)
    dL_dmeans3D = torch.zeros((P, 3), **o)
    dL_dmeans2D = torch.zeros((P, 3), **o)
    dL_dcolors = torch.zeros((P, 3), **o)
    dL_ddepths = torch.zeros((P, 1), **o)
    dL_dconic = torch.zeros((P, 2, 2), **o)
    dL_dopacity = torch.zeros((P, 1), **o)
    dL_dcov3D = torch.zeros((P, 6), **o)
    dL_dsh = torch.zeros((P, M, 3), **o)
    dL_dscales = torch.zeros((P, 3), **o)
    dL_drotations = torch.zeros((P, 4), **o)
    dL_dflow = torch.zeros((P, 3), **o)
    dL_dsem = torch.zeros((P, D_S), **o)
    keep = [_c(t) for t in (background, means3D, sh, colors, flow_points, semantic, scales, rotations, cov3D_precomp,
                            viewmatrix, projmatrix, campos, radii, dL_dout_color, dL_dout_depth, dL_dout_flow,
                            dL_dout_semantic, grad_img_opacity, img_opacity)]
    (background, means3D, sh, colors, flow_points, semantic, scales, rotations, cov3D_precomp, viewmatrix, projmatrix,
     campos, radii, dL_dout_color, dL_dout_depth, dL_dout_flow, dL_dout_semantic, grad_img_opacity,
     img_opacity) = keep
    if P != 0:
        _join_default_stream(dev)
        with torch.cuda.device(dev):
            st = lib.ref_backward(
                C.c_int(P), C.c_int(int(degree)), C.c_int(M), C.c_int(int(R)), C.c_int(D_S), _p(background),
                C.c_int(W), C.c_int(H), _p(means3D), _p(sh), _p(colors), _p(flow_points), _p(semantic), _p(scales),
                C.c_float(scale_modifier), _p(rotations), _p(cov3D_precomp), _p(viewmatrix), _p(projmatrix),
                _p(campos), C.c_float(tan_fovx), C.c_float(tan_fovy), _p(radii), _p(geomBuffer), _p(binningBuffer),
                _p(imageBuffer), _p(dL_dout_color), _p(dL_dout_depth), _p(dL_dout_flow), _p(dL_dout_semantic),
                _p(dL_dmeans2D), _p(dL_dconic), _p(dL_dopacity), _p(dL_dcolors), _p(dL_ddepths), _p(dL_dmeans3D),
                _p(dL_dcov3D), _p(dL_dsh), _p(dL_dscales), _p(dL_drotations), _p(dL_dflow), _p(dL_dsem),
                _p(grad_img_opacity), _p(img_opacity), C.c_int(int(inv_depth)), C.c_int(int(debug)))
        _join_default_stream(dev)
        if st != 0:
            raise RuntimeError("reference backward failed")
    return (dL_dmeans2D, dL_dcolors, dL_dopacity, dL_dmeans3D, dL_dcov3D, dL_dsh, dL_dscales, dL_drotations,
            dL_dflow, dL_dsem)


def mark_visible(means3D, viewmatrix, projmatrix):
    lib = load()
    P = means3D.shape[0]
    present = torch.zeros((P,), dtype=torch.bool, device=means3D.device)
    if P:
        torch.cuda.synchronize()
        lib.ref_mark_visible(C.c_int(P), _p(means3D.contiguous()), _p(viewmatrix.contiguous()),
                             _p(projmatrix.contiguous()), _p(present))
        torch.cuda.synchronize()
    return present


def dist_cuda2(points):
    lib = load()
    P = points.shape[0]
    out = torch.zeros((P,), dtype=torch.float32, device=points.device)
    torch.cuda.synchronize()
    lib.ref_dist_cuda2(C.c_int(P), _p(points.contiguous()), _p(out))
    torch.cuda.synchronize()
    return out


def _view(buf, off, count, dtype):
    esz = torch.empty((), dtype=dtype).element_size()
    return buf[off:off + count * esz].view(dtype)


def inspect(geomBuffer, binningBuffer, imgBuffer, P, R, W, H):
    """Re-interpret the three reference arenas (rasterizer_impl.cu:155-194) as named tensors."""
    lib = load()
    res = {}
    if geomBuffer.numel():
        gl = GeomLayout()
        lib.ref_geom_offsets(C.c_void_p(geomBuffer.data_ptr()), C.c_size_t(P), C.byref(gl))
        res["depths"] = _view(geomBuffer, gl.depths, P, torch.float32)
        res["clamped"] = _view(geomBuffer, gl.clamped, 3 * P, torch.bool).view(P, 3)
        res["means2D"] = _view(geomBuffer, gl.means2D, 2 * P, torch.float32).view(P, 2)
        res["cov3D"] = _view(geomBuffer, gl.cov3D, 6 * P, torch.float32).view(P, 6)
        res["conic_opacity"] = _view(geomBuffer, gl.conic_opacity, 4 * P, torch.float32).view(P, 4)
        res["rgb"] = _view(geomBuffer, gl.rgb, 3 * P, torch.float32).view(P, 3)
        res["tiles_touched"] = _view(geomBuffer, gl.tiles_touched, P, torch.int32)
        res["point_offsets"] = _view(geomBuffer, gl.point_offsets, P, torch.int32)
    if binningBuffer.numel() and R > 0:
        bl = BinningLayout()
        lib.ref_binning_offsets(C.c_void_p(binningBuffer.data_ptr()), C.c_size_t(R), C.byref(bl))
        res["point_list"] = _view(binningBuffer, bl.point_list, R, torch.int32)
        res["point_list_keys"] = _view(binningBuffer, bl.point_list_keys, R, torch.int64)
    if imgBuffer.numel():
        il = ImageLayout()
        lib.ref_image_offsets(C.c_void_p(imgBuffer.data_ptr()), C.c_size_t(W * H), C.byref(il))
        res["n_contrib"] = _view(imgBuffer, il.n_contrib, W * H, torch.int32).view(H, W)
        tiles = ((W + 15) // 16) * ((H + 15) // 16)
        res["ranges"] = _view(imgBuffer, il.ranges, 2 * tiles, torch.int32).view(tiles, 2)
    return res
