"""Stub: reanalyze.py only names fbx.prioritised_flat_buffer.ExperiencePair in an annotation."""


class prioritised_flat_buffer:  # noqa: N801
    ExperiencePair = object
