"""Host logic of the design-tail fusion (mrphy._ops.tag_design / _design_of): which waveforms may be differentiated w.r.t. their
design variables inside the simulation's gradient epilogue.  CPU tensors only -- no kernel is launched."""
import torch

from mrphy import _ops


def _chain():
    rho = torch.randn(2, 1, 9, requires_grad=True)
    theta = torch.randn(2, 1, 9, requires_grad=True)
    ts = torch.randn(2, 3, 9, requires_grad=True)
    rfmax, smax, dt = torch.tensor([0.2]), torch.tensor([[1e4, 1e4, 1e4]]), torch.tensor([4e-6])
    rf = torch.cat([rho.cos(), theta.sin()], dim=1) * 0.1          # stand-ins for the kernel's outputs (non-leaf, require grad)
    gr = ts.cumsum(dim=2) * 1e-2
    _ops.tag_design(rf, gr, rho, theta, rfmax, ts, smax, dt, 1, 1)
    return rho, theta, ts, rf, gr


def test_tagged_outputs_are_recognised_only_as_the_same_object():
    rho, theta, ts, rf, gr = _chain()
    r = _ops._design_of(rf, rf)
    assert r is not None and r.kind == 1 and r.tensors[0] is rho and r.tensors[1] is theta
    g = _ops._design_of(gr, gr)
    assert g is not None and g.kind == 1 and g.tensors[0] is ts
    assert _ops._design_of(rf, rf.to(torch.float64)) is None       # a cast / copy is another tensor: plain path
    assert _ops._design_of(rf * 1.0, rf * 1.0) is None             # no record on derived tensors


def test_in_place_edits_invalidate_the_record():
    rho, theta, ts, rf, gr = _chain()
    with torch.no_grad():
        rf.mul_(2.0)                                               # the waveform no longer is chain(rho, theta)
    assert _ops._design_of(rf, rf) is None
    assert _ops._design_of(gr, gr) is not None
    with torch.no_grad():
        ts.add_(1.0)                                               # nor is gr the chain of the CURRENT ts
    assert _ops._design_of(gr, gr) is None


def test_callers_who_want_the_waveform_gradient_keep_the_two_stage_path():
    rho, theta, ts, rf, gr = _chain()
    rf.retain_grad()
    assert _ops._design_of(rf, rf) is None
    gr.register_hook(lambda g: g)
    assert _ops._design_of(gr, gr) is None
    rho2, theta2, ts2, rf2, gr2 = _chain()
    assert _ops._design_of(rf2.detach(), rf2.detach()) is None     # nothing to differentiate


def test_slew_output_is_not_tagged_as_a_gradient():
    ts = torch.randn(1, 3, 5, requires_grad=True)
    s = ts.atan()
    _ops.tag_design(None, s, None, None, None, ts, torch.ones(1, 3), None, 0, 3)      # gr_kind 3: ts -> s, not a gradient
    assert getattr(s, '_mrphy_design', None) is None
