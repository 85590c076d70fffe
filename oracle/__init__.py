"""CPU oracle for the proFit Custom-GP hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is product code.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it, and only as the checker
or as the timed CPU baseline -- never as a fallback for the CUDA path.
"""
from .gp_oracle import *  # noqa: F401,F403
