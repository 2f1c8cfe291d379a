"""TEST INFRASTRUCTURE — the seeded synthetic weights / inputs live in cogview_b200.recipes (a neutral module that
bench.py's product arm and smoke() may use without touching oracle/); re-exported here for the tests and the
golden-vector generator."""
from cogview_b200.recipes import *  # noqa: F401,F403
from cogview_b200.recipes import CONFIG1, COGVIEW_4B, IMG_VOCAB  # noqa: F401
