"""Per-frame error of the GPU encoder vs the bf16-emulating oracle (diagnostic)."""
import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import reverb_b200
from reverb_b200 import synth
from oracle import model_ref, pipeline_ref

for causal, norm, seed in [(True, "layer_norm", 0), (False, "batch_norm", 1)]:
    d = tempfile.mkdtemp()
    synth.write_model_dir(d, causal=causal, cnn_module_norm=norm, seed=seed, blank_rate=0.5)
    wav = synth.write_wav(os.path.join(d, "a.wav"), synth.synth_audio(11.3, seed=1234))
    m = reverb_b200.load_model(d)
    orc = pipeline_ref.OracleASR(d)
    feats = orc.compute_feats(wav)
    cat = torch.tensor([0.7, 0.3])
    fb, fl = next(orc.feats_batcher(feats, 400, 2))
    model_ref.EMULATE_BF16 = True
    with torch.no_grad():
        want, lens, _ = orc.forward_encoder(fb, fl, cat)
        wlogp = model_ref.ctc_logprobs(want, orc.sd)
    model_ref.EMULATE_BF16 = False
    enc, enc_lens = m.model._forward_encoder(fb.cuda(), fl, cat)
    logp = m.model.ctc_logprobs(enc).cpu()
    got = enc.cpu()
    for b in range(fb.shape[0]):
        n = int(enc_lens[b])
        err = (got[b, :n] - want[b, :n]).abs().amax(-1)
        rms = ((got[b, :n] - want[b, :n]) ** 2).mean(-1).sqrt() / (want[b, :n] ** 2).mean(-1).sqrt()
        lerr = (logp[b, :n] - wlogp[b, :n]).abs()
        lerr[wlogp[b, :n] < -12] = 0
        le = lerr.amax(-1)
        top = torch.topk(le, 5)
        print(f"causal={causal} b={b} n={n} enc max-abs-err/frame: mean {err.mean():.4f} max {err.max():.4f} at t={int(err.argmax())}; "
              f"rel-rms/frame mean {rms.mean():.2e} max {rms.max():.2e}")
        print("   logp err top frames:", [(int(i), round(float(v), 3)) for v, i in zip(top.values, top.indices)])
        print("   per-frame rel-rms first 8:", [f"{x:.1e}" for x in rms[:8].tolist()], "last 4:", [f"{x:.1e}" for x in rms[-4:].tolist()])
