"""Alias of the `reverb` console script entry point (`wenet.bin.recognize_wav:main`)."""
from reverb_b200.recognize_wav import get_args, main  # noqa: F401

if __name__ == "__main__":
    main()
