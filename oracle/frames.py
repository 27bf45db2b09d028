"""Seeded synthetic frames: the generator lives in prisma_b200/synthetic.py (bench.py and the band scripts' offline runs use
it without touching oracle/); the oracle tools and the tests import it from here."""
from prisma_b200.synthetic import synthetic_frame  # noqa: F401
