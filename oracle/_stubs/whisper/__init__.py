# Stub of the absent `openai-whisper` package: the reference imports
# whisper.tokenizer.LANGUAGES unconditionally (asr/wenet/utils/common.py:23).
# Test infrastructure only -- never imported by the product path.
