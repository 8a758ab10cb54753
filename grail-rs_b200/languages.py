"""languages::generic() (reference src/languages/mod.rs:4-34)"""
from .text import Language, generic_language

generic = generic_language
__all__ = ["generic", "Language"]
