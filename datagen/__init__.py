"""Input production for tests and benchmarks (NOT part of the product path).

`zsyn` is the deterministic "zsyn-v1" corpus of SURVEY.md Appendix D; `refwriter` drives the reference
CPU writer (oracle/_ref/libzseek_ref.so) to turn a corpus into seekable files, as BASELINE.json's
north_star prescribes ("the write path stays the reference CPU writer and is used only to produce the
inputs").
"""
