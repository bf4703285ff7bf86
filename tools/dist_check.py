"""Multi-GPU check (run under torchrun, one rank per GPU): chunk-sharded long-form decode with ONE NCCL all-gather
must reproduce the single-GPU sequential result exactly, on every rank."""
import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import reverb_b200
from reverb_b200 import synth
from reverb_b200.dist import transcribe_sharded
from reverb_b200.reverb import get_output

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
d = os.path.join(tempfile.gettempdir(), "rvb_dist_model")
if rank == 0:
    synth.write_model_dir(d, seed=5, blank_rate=0.5)
dist.barrier()
asr = reverb_b200.load_model(d, gpu=local)
pcm = synth.synth_audio(47.3, seed=99)            # 4728 frames -> 8 chunks of 600 (+ a padded tail chunk)
kw = dict(mode="attention_rescoring", chunk_size=600, batch_size=3, beam_size=10, ctc_weight=0.1, reverse_weight=0.3,
          verbatimicity=0.7)
hyps = transcribe_sharded(asr, pcm, **kw)
ctm = get_output("ctm", asr.tokenizer, "dist.wav", hyps, 230, 600, 10, 40)
# single-GPU reference on this rank: the public API on the same audio
wav = synth.write_wav(os.path.join(tempfile.gettempdir(), f"rvb_dist_{rank}.wav"), pcm)
ref = asr.transcribe(wav, mode="attention_rescoring", format="ctm", chunk_size=600, batch_size=3, beam_size=10,
                     ctc_weight=0.1, reverse_weight=0.3, verbatimicity=0.7)
ref = ref.replace(os.path.basename(wav), "dist.wav")
assert len(hyps) == 8, len(hyps)
assert ctm == ref, f"rank {rank}: sharded result differs from the sequential one"
gathered = [None] * world
dist.all_gather_object(gathered, ctm)
assert all(g == ctm for g in gathered)
if rank == 0:
    print(f"DIST_OK world={world} chunks={len(hyps)} ctm_lines={ctm.count(chr(10)) + 1}")
dist.destroy_process_group()
