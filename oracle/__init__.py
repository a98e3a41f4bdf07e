"""CPU parity oracle for the bskit FFT-bispectrum hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``bskit_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and the CPU legs of
``bench.py`` (``cpu_baseline`` and ``--impl reference``) use it, and there only
as the checker / the timed CPU baseline, never as the product path.
"""
