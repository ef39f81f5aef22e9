"""csrc/optim.cu (gradient norm, clip + Adam) executed through the CPU emulation shim against torch."""
import ctypes as C

import pytest
import torch

import emu_harness as H


@pytest.mark.parametrize("clip", [0.0, 1.5])
def test_emulated_adam_matches_torch(clip):
    lib = H.load_emu()
    fptr = C.c_void_p
    lib.tdanet_grad_sqnorm.argtypes = [fptr, C.c_size_t, fptr, fptr]
    lib.tdanet_adam_step.argtypes = [fptr, fptr, fptr, fptr, C.c_size_t, C.c_float, C.c_float, C.c_float, C.c_float,
                                     C.c_float, C.c_float, fptr, fptr, fptr]
    g = torch.Generator().manual_seed(5)
    n = 4099
    p0 = torch.randn(n, generator=g)
    ref_p = torch.nn.Parameter(p0.clone().double())
    opt = torch.optim.Adam([ref_p], lr=1e-3, betas=(0.9, 0.999), eps=1e-8)
    p, m, v = p0.clone(), torch.zeros(n), torch.zeros(n)
    step = torch.zeros(1, dtype=torch.int32)
    sq = torch.zeros(2, dtype=torch.float64)
    world = 2   # gradients arrive summed over two ranks: the kernel folds the 1/world in
    for it in range(4):
        grad = torch.randn(n, generator=g) * (10.0 ** (it - 1))
        ref_p.grad = grad.clone().double()
        if clip > 0:
            torch.nn.utils.clip_grad_norm_([ref_p], clip)
        opt.step()
        summed = (grad * world).contiguous()
        assert lib.tdanet_grad_sqnorm(summed.data_ptr(), n, sq.data_ptr(), None) == 0
        assert lib.tdanet_adam_step(p.data_ptr(), summed.data_ptr(), m.data_ptr(), v.data_ptr(), n, 1e-3, 0.9, 0.999,
                                    1e-8, clip, 1.0 / world, sq.data_ptr() if clip > 0 else None, step.data_ptr(),
                                    None) == 0
    assert step.item() == 4
    assert (p.double() - ref_p.detach()).abs().max().item() < 2e-6
