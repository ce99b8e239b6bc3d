"""CPU oracle for the BYOL training-step hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``byol_b200/`` may import this package: only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` do, and only as the checker / reported baseline.
"""
