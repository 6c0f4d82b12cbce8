"""GPU: pin the tcgen05 shared-memory descriptor semantics the fused conv kernels rely on -- a convolution tap is a
ROW-SHIFTED view of one activation tile, i.e. the A descriptor's start address moves by whole rows."""
import pytest
import torch

from erfnet_pytorch_b200 import ops

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("layout", [0, 2])
@pytest.mark.parametrize("n", [64, 128, 16])
@pytest.mark.parametrize("shift", [0, 8, 1, 3, 34, 129])
def test_row_shifted_descriptor(layout, n, shift):
    gen = torch.Generator().manual_seed(shift * 7 + n)
    rows, k = 128 + 136, 64
    a = (torch.randn(rows, k, generator=gen)).bfloat16()
    b = (torch.randn(n, k, generator=gen)).bfloat16()
    d, _ = ops.umma_probe(a.cuda(), b.cuda(), shift, layout)
    ref = a[shift:shift + 128].float() @ b.float().t()
    torch.testing.assert_close(d.cpu(), ref, rtol=1e-3, atol=1e-2)


def test_mma_issue_rate():
    """Informational: cycles per 128xNx16 MMA for both layouts (printed with -s)."""
    for layout in (0, 2):
        for n in (64, 128):
            a = torch.randn(256, 64).bfloat16().cuda()
            b = torch.randn(n, 64).bfloat16().cuda()
            _, cyc = ops.umma_probe(a, b, 1, layout, iters=256)
            torch.cuda.synchronize()
            print(f"layout={layout} N={n}: {cyc.item() / (256 * 4):.1f} cycles per MMA (128x{n}x16)")
