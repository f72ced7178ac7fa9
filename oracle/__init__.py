"""CPU oracle for the scoring hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this package.  The product package
(``mmf_b200``) never imports it and has no CPU fallback.

Parity pin: the reference ships no tests or golden vectors, so the pin is
``tests/golden/*.npz`` -- outputs of the reference's OWN methods
(``/root/reference/misinfo_forensics.py``, ``clip_similarity_engine.py``) run
in the build container by ``tests/golden/make_golden.py``; the oracle is checked
against those fixtures by ``tests/test_oracle_golden.py``.
"""
from .reference_port import *  # noqa: F401,F403
