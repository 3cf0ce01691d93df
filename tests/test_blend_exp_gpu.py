"""The blend kernels evaluate exp(power) for two splats at once with packed FP32x2 instructions
(adgs_b200/csrc/blend.cu:exp_pair). n_contrib / img_opacity are only bit-exact against the reference's
renderCUDA (RZ/cuda_rasterizer/forward.cu:345) if that exponential equals expf() bit for bit, so the
library carries a device self-test over EVERY float of the domain."""
import pytest
import torch

from adgs_b200 import _lib as L

pytestmark = pytest.mark.gpu


def test_exp_pair_is_bit_identical_to_expf_on_its_whole_domain():
    lib = L.load()
    out = torch.zeros(2, dtype=torch.int64, device="cuda")
    L.check(lib.adgs_selftest_exp_pair(out.data_ptr(), torch.cuda.current_stream().cuda_stream), "selftest_exp_pair")
    torch.cuda.synchronize()
    bad, checked = int(out[0]), int(out[1])
    # [-87, 87] holds 2 * (0x42AE0000 + 1) bit patterns (both signs of zero, denormals included); NaNs: 2 * (2^23 - 1)
    assert checked == 2 * (0x42AE0000 + 1) + 2 * (2 ** 23 - 1)
    assert bad == 0, f"{bad} of {checked} arguments differ from expf()"
