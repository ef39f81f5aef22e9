"""Checkpoint interchange (host logic, CPU): flat Adam moments <-> torch.optim.Adam.state_dict(), the Lightning
`audio_model.` prefix, and the plateau scheduler against torch's."""
import torch

import tdanet_b200.look2hear as look2hear
from tdanet_b200.look2hear.system import checkpoint as ck


def _tiny():
    torch.manual_seed(0)
    return look2hear.models.TDANetBest(out_channels=16, in_channels=32, num_blocks=2, upsampling_depth=4,
                                       enc_kernel_size=4, num_sources=2, sample_rate=8000)


def test_adam_state_round_trip_through_torch_optimizer():
    m = _tiny()
    named = [(n, p.shape) for n, p in m.named_parameters()]
    offs, total = ck.param_layout(named)
    g = torch.Generator().manual_seed(1)
    exp_avg, exp_avg_sq = torch.randn(total, generator=g), torch.rand(total, generator=g)
    sd = ck.adam_state_to_torch(named, exp_avg, exp_avg_sq, step=7, lr=5e-4)
    opt = torch.optim.Adam(m.parameters(), lr=1e-3)
    opt.load_state_dict(sd)                              # the reference trainer can resume from it
    assert opt.param_groups[0]["lr"] == 5e-4
    for (n, p), o in zip(m.named_parameters(), offs):
        st = opt.state[p]
        assert int(st["step"]) == 7
        assert torch.equal(st["exp_avg"].reshape(-1), exp_avg[o:o + p.numel()])
        assert torch.equal(st["exp_avg_sq"].reshape(-1), exp_avg_sq[o:o + p.numel()])
    a2, b2 = torch.empty(total), torch.empty(total)
    step, lr, betas, eps = ck.torch_to_adam_state(named, opt.state_dict(), a2, b2)
    assert (step, lr, betas, eps) == (7, 5e-4, (0.9, 0.999), 1e-8)
    for (n, p), o in zip(m.named_parameters(), offs):
        assert torch.equal(a2[o:o + p.numel()], exp_avg[o:o + p.numel()])
        assert torch.equal(b2[o:o + p.numel()], exp_avg_sq[o:o + p.numel()])


def test_lightning_prefix_round_trip(tmp_path):
    m = _tiny()
    ckpt = {"state_dict": ck.lightning_state_dict(m.state_dict())}
    assert all(k.startswith("audio_model.") for k in ckpt["state_dict"])
    path = tmp_path / "epoch=1.ckpt"
    torch.save(ckpt, path)
    m2 = look2hear.models.TDANetBest.from_pretrain("TDANetBest", str(path), out_channels=16, in_channels=32, num_blocks=2,
                                                   upsampling_depth=4, enc_kernel_size=4, num_sources=2, sample_rate=8000)
    for (k, v), (k2, v2) in zip(m.state_dict().items(), m2.state_dict().items()):
        assert k == k2 and torch.equal(v, v2)
    assert set(ck.strip_lightning_prefix(ckpt["state_dict"])) == set(m.state_dict())


def test_plateau_scheduler_matches_torch():
    class Holder:
        lr = 1e-3
    ours_opt = Holder()
    ours = look2hear.system.ReduceLROnPlateau(ours_opt, factor=0.5, patience=2)
    p = torch.nn.Parameter(torch.zeros(1))
    ref_opt = torch.optim.Adam([p], lr=1e-3)
    ref = torch.optim.lr_scheduler.ReduceLROnPlateau(ref_opt, factor=0.5, patience=2)
    for v in [5.0, 4.0, 4.0, 4.0, 4.0, 3.9, 3.9, 3.9, 3.9, 3.9, 3.9, 3.9, 1.0, 1.0]:
        ours.step(v)
        ref.step(v)
        assert abs(ours_opt.lr - ref_opt.param_groups[0]["lr"]) < 1e-12
    st = ours.state_dict()
    assert st["best"] == ref.state_dict()["best"] and st["num_bad_epochs"] == ref.state_dict()["num_bad_epochs"]
