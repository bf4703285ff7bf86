"""Speaker-attribution leg of the reference (diarization/): RTTM I/O, word -> speaker assignment and the STM writer
(assign_words2speakers.py), host side.  The neural forward (pyannote pipeline behind infer_pyannote3.0.py) lives in
reverb_b200/diarization/segmentation.py (+ csrc/diar.cu)."""
from .rttm import Turn, load_rttm, write_rttm  # noqa: F401
from .words2speakers import SpeakerIndex, assign_words_to_speakers, read_ctm, speaker_for_segment, write_stm  # noqa: F401
