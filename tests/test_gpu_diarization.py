"""Diarization forward on the GPU (csrc/diar_seg.cu through include/rvb_diar.h) vs the torch restatement in
oracle/diar_ref.py on the same synthetic weights and audio.  ** parity unpinned ** against pyannote itself (absent
offline); what is pinned here is kernel == stock torch modules for the published architecture.

Tolerances: everything is fp32 with a different summation order, the LSTM recurrence runs 589 steps: SincNet output
2e-4 abs (values O(1)), log-probabilities 2e-3 abs, arg-max class identical except at near-ties (< 1e-3 margin)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def seg():
    from reverb_b200.diarization import synth
    from reverb_b200.diarization.segmentation import SegmentationModel
    sd = synth.segmentation_state_dict(0)
    return sd, SegmentationModel(sd)


def _windows(n, seconds=10.0):
    from reverb_b200.diarization import synth
    return np.stack([synth.synthetic_speech(seconds, seed=10 + i, turns=2 + i % 2) for i in range(n)])


def test_segmentation_forward_vs_torch_oracle(seg):
    from oracle import diar_ref
    sd, model = seg
    wav = _windows(3)
    ref_net = diar_ref.PyanNetRef(sd)
    w = torch.from_numpy(wav)
    ref_sinc = ref_net.sincnet(w).transpose(1, 2).numpy()
    ref = ref_net(w).numpy()
    got, sinc = model.forward(w.cuda(), return_sincnet=True)
    got, sinc = got.cpu().numpy(), sinc.cpu().numpy()
    assert got.shape == ref.shape == (3, 589, 7)
    print("sincnet max abs diff", np.abs(sinc - ref_sinc).max(), "logp max abs diff", np.abs(got - ref).max())
    assert np.abs(sinc - ref_sinc).max() < 2e-4
    assert np.abs(got - ref).max() < 2e-3
    am, rm = got.argmax(-1), ref.argmax(-1)
    bad = np.argwhere(am != rm)
    for b, t in bad:
        top2 = np.sort(ref[b, t])[-2:]
        assert top2[1] - top2[0] < 1e-3, f"class differs at ({b},{t}) beyond a near-tie"
    print("argmax agreement", (am == rm).mean(), "classes used", sorted(set(rm.flatten().tolist())))


def test_segmentation_batch_tail_and_short_window(seg):
    """batch sizes that do not fill the 8-window LSTM tile, and a 5 s window (293 frames)"""
    from oracle import diar_ref
    sd, model = seg
    ref_net = diar_ref.PyanNetRef(sd)
    wav = _windows(9, seconds=5.0)
    w = torch.from_numpy(wav)
    got = model.forward(w.cuda()).cpu().numpy()
    assert got.shape == (9, 293, 7)
    ref = ref_net(w[[0, 7, 8]]).numpy()
    assert np.abs(got[[0, 7, 8]] - ref).max() < 2e-3
    # a window's result does not depend on its batch neighbours
    solo = model.forward(w[8:9].cuda()).cpu().numpy()
    assert np.array_equal(solo[0], got[8])


def test_segmentation_rejects_bad_input(seg):
    _, model = seg
    with pytest.raises(ValueError):
        model.forward(torch.zeros(1, 100, device="cuda"))
