"""Scorer models with the reference's module layout (models/*.py)."""
