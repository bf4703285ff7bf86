"""Host half of the speaker-attribution leg (reference: diarization/assign_words2speakers.py:24-87).  The reference's
own dependencies (intervaltree, pyannote) are absent, so the cases below are worked out by hand from the reference
code (each comment says which branch of speaker_for_segment it exercises) — "parity unpinned" by a live run."""
import io

import pytest

from reverb_b200.diarization import (SpeakerIndex, Turn, assign_words_to_speakers, load_rttm, speaker_for_segment,
                                     write_rttm, write_stm)

TURNS = [(0.0, 2.0, "A"), (1.5, 4.0, "B"), (6.0, 7.0, "A"), (6.5, 9.0, "C")]


def test_speaker_for_segment_branches():
    tree = SpeakerIndex(TURNS)
    assert speaker_for_segment(0.2, 0.5, tree) == "A"            # :37-38 exactly one overlapping turn
    assert speaker_for_segment(2.5, 1.0, tree) == "B"
    assert speaker_for_segment(1.4, 0.4, tree) == "A"            # :51-61 overlap: A 0.4, B 0.3 -> majority A
    assert speaker_for_segment(1.6, 1.0, tree) == "B"            # A 0.4, B 1.0
    assert speaker_for_segment(6.2, 2.0, tree) == "C"            # A 0.8 (one turn), C 1.7
    assert speaker_for_segment(4.5, 0.5, tree) == "B"            # :42-49 no overlap: gap to B 0.5, to A(6.0) 1.0
    assert speaker_for_segment(5.4, 0.3, tree) == "A"            # gap to A(6.0) 0.3, to B 1.4
    assert speaker_for_segment(20.0, 1.0, tree) == "C"           # after the last turn
    # zero-duration word: tree[a:a] is empty -> nearest.  At 2.0 both A (0, 2) [gap 2 - 2] and B (1.5, 4) [overlap]
    # are at distance 0: an exact tie, which the reference resolves by set order; here: the earliest turn
    assert speaker_for_segment(2.0, 0.0, tree) == "A"
    assert speaker_for_segment(2.1, 0.0, tree) == "B"
    assert speaker_for_segment(1.0, 1.0, SpeakerIndex([])) == ""  # :47-48 empty tree
    # touching is not overlapping (half-open): a word ending exactly where a turn starts
    assert speaker_for_segment(5.0, 1.0, tree) == "A"            # [5, 6) vs A [6, 7): no overlap, distance 0 -> A
    # same speaker in several overlapping turns: overlaps add up (defaultdict(int) :55-59)
    t2 = SpeakerIndex([(0.0, 1.0, "A"), (2.0, 3.0, "A"), (0.5, 2.6, "B")])
    assert speaker_for_segment(0.0, 3.0, t2) == "B"              # A 1.0 + 1.0 = 2.0 < B 2.1
    assert speaker_for_segment(0.0, 2.5, t2) == "B"              # A 1.0 + 0.5 = 1.5 < B 2.0
    assert speaker_for_segment(0.0, 1.4, t2) == "A"              # A 1.0 > B 0.9


def test_null_turns_are_rejected_like_intervaltree():
    with pytest.raises(ValueError):
        SpeakerIndex([(1.0, 1.0, "A")])


def test_rttm_roundtrip_and_stm(tmp_path):
    turns = [Turn(b, e, lab) for b, e, lab in TURNS]
    buf = io.StringIO()
    write_rttm(buf, "call1", reversed(turns))
    text = buf.getvalue()
    assert text.splitlines()[0] == "SPEAKER call1 1 0.000 2.000 <NA> <NA> A <NA> <NA>"
    assert text.splitlines()[1] == "SPEAKER call1 1 1.500 2.500 <NA> <NA> B <NA> <NA>"
    rttm = tmp_path / "call1.rttm"
    rttm.write_text(text + "SPKR-INFO call1 1 <NA> <NA> <NA> unknown A <NA> <NA>\n")
    back = load_rttm(str(rttm))
    assert list(back) == ["call1"] and back["call1"] == turns
    ctm = tmp_path / "call1.ctm"
    ctm.write_text("call1.wav 0 0.20 0.50 hello 0.98\ncall1.wav 0 1.60 1.00 there 0.71\ncall1.wav 0 4.50 0.50 <sp5> 0.00\n")
    stm = tmp_path / "call1.stm"
    write_stm(str(rttm), str(ctm), str(stm))
    assert stm.read_text() == ("call1 1 A 0.200 0.700 hello\n"
                               "call1 1 B 1.600 2.600 there\n"
                               "call1 1 B 4.500 5.000 <sp5>\n")
    rows = [ln.split(" ") for ln in ctm.read_text().splitlines()]
    assert assign_words_to_speakers(rows, turns, "x")[0] == "x 1 A 0.200 0.700 hello"


def test_two_uris_are_an_error(tmp_path):
    rttm = tmp_path / "two.rttm"
    rttm.write_text("SPEAKER a 1 0.000 1.000 <NA> <NA> S <NA> <NA>\nSPEAKER b 1 0.000 1.000 <NA> <NA> S <NA> <NA>\n")
    (tmp_path / "w.ctm").write_text("a 0 0.0 0.5 w 1.0\n")
    with pytest.raises(AssertionError):
        write_stm(str(rttm), str(tmp_path / "w.ctm"), str(tmp_path / "o.stm"))
