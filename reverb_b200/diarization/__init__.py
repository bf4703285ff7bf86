"""Speaker-attribution leg of the reference (diarization/): RTTM I/O, word -> speaker assignment and the STM writer
(assign_words2speakers.py) on the host; the neural forward of the pyannote pipeline behind infer_pyannote3.0.py on the GPU:
segmentation.py (csrc/diar_seg.cu), embedding.py (csrc/diar_emb.cu), pipeline.py (windowing / clustering glue), infer.py
(CLI).  The GPU modules are imported lazily: `from reverb_b200.diarization.pipeline import SpeakerDiarization`."""
from .rttm import Turn, load_rttm, write_rttm  # noqa: F401
from .words2speakers import SpeakerIndex, assign_words_to_speakers, read_ctm, speaker_for_segment, write_stm  # noqa: F401
