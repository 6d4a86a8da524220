"""smoqyelph_b200 -- B200-native hot path of SmoQyElPhQMC.jl behind the reference's operator API.

Only `model` (pure numpy tables) is imported eagerly; `lib` / `api` load the CUDA C-ABI library
and fail loudly if it is missing (there is no CPU fallback).
"""
from . import model  # noqa: F401
