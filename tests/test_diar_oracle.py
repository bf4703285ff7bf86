"""Host-side checks of the diarization oracle restatement and the synthetic weights (no GPU).

** parity unpinned ** (oracle/diar_ref.py): these tests pin the restatement's internal consistency and the published
shape constants (10 s -> 589 frames, 7 powerset classes), not the shipped checkpoints."""
import numpy as np
import torch

from oracle import diar_ref
from reverb_b200.diarization import synth
from reverb_b200.diarization.segmentation import powerset_mapping, powerset_to_multilabel


def test_frame_arithmetic_matches_the_published_589_frames_per_10s():
    assert diar_ref.seg_num_frames(160000) == 589
    assert diar_ref.seg_num_frames(80000) == 293


def test_sinc_filter_bank_is_a_mirrored_bandpass_bank():
    sd = synth.segmentation_state_dict(0)
    f = diar_ref.sinc_filters(torch.from_numpy(sd["sincnet.conv1d.0.filterbank.low_hz_"]),
                              torch.from_numpy(sd["sincnet.conv1d.0.filterbank.band_hz_"]))
    assert f.shape == (80, 251)
    cos, sin = f[:40], f[40:]
    assert torch.allclose(cos, torch.flip(cos, dims=[1]))            # even
    assert torch.allclose(sin, -torch.flip(sin, dims=[1]))           # odd
    assert torch.allclose(cos[:, 125], torch.ones(40))
    # the pass band of filter c contains its centre frequency: response there >> response far outside
    low = 50 + np.abs(sd["sincnet.conv1d.0.filterbank.low_hz_"][:, 0])
    band = 50 + np.abs(sd["sincnet.conv1d.0.filterbank.band_hz_"][:, 0])
    t = torch.arange(251, dtype=torch.float32) / 16000
    for c in (5, 20, 35):
        fc = float(low[c] + band[c] / 2)
        inside = (cos[c] * torch.cos(2 * np.pi * fc * t)).sum().abs()
        outside = (cos[c] * torch.cos(2 * np.pi * (fc + 6 * band[c] + 600) * t)).sum().abs()
        assert inside > 5 * outside


def test_pyannet_oracle_shapes_and_normalisation():
    sd = synth.segmentation_state_dict(1)
    net = diar_ref.PyanNetRef(sd)
    wav = torch.from_numpy(synth.synthetic_speech(2.0, seed=3, turns=2)).view(1, -1)
    logp = net(wav)
    assert logp.shape == (1, diar_ref.seg_num_frames(wav.shape[1]), 7)
    assert torch.allclose(logp.exp().sum(-1), torch.ones(1, logp.shape[1]), atol=1e-5)
    # the synthetic classifier is not degenerate: more than one class wins somewhere
    assert len(set(logp.argmax(-1).flatten().tolist())) > 1


def test_powerset_mapping_order():
    m = powerset_mapping(3, 2)
    assert m.tolist() == [[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1], [1, 1, 0], [1, 0, 1], [0, 1, 1]]
    logp = torch.log(torch.tensor([[[0.1, 0.6, 0.1, 0.05, 0.05, 0.05, 0.05], [0.0, 0.0, 0.0, 0.0, 0.1, 0.1, 0.8]]]) + 1e-9)
    ml = powerset_to_multilabel(logp, torch.from_numpy(m))
    assert ml.tolist() == [[[1, 0, 0], [0, 1, 1]]]


def test_stats_pool_weighted_equals_unweighted_for_unit_weights():
    x = torch.randn(2, 6, 40, generator=torch.Generator().manual_seed(0))
    a = diar_ref.stats_pool(x, None)
    b = diar_ref.stats_pool(x, torch.ones(2, 40))
    assert torch.allclose(a, b, atol=1e-5)
    # weights given at another frame rate are nearest-interpolated
    c = diar_ref.stats_pool(x, torch.ones(2, 13))
    assert torch.allclose(a, c, atol=1e-5)
