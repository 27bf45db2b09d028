"""Seeded test weights: the generator lives in prisma_b200/seeded_weights.py (the band scripts expose it as
--seeded-weights for offline runs); the oracle and the tests import it from here."""
from prisma_b200.seeded_weights import (DA_CONFIGS, MIDAS_CONFIGS, SOLO_CONFIGS, ZOE_CONFIG, make_da_weights,  # noqa: F401
                                        make_zoe_weights,
                                        make_midas_weights, make_raft_weights, make_solo_weights)
