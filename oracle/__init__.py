"""CPU oracle for the YOLO detection post-processing path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``ultralytics_pro_b200/`` may import this
package.  The only legitimate importers are ``tests/``, ``__graft_entry__.smoke()``
and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``.

Parity status: the reference (Chriz122/ultralytics_pro) ships no tests, golden
vectors or known-answer fixtures for this path.  The oracle is
therefore pinned by *running the reference itself* in the build container
(``oracle/make_golden.py`` imports it from ``/root/reference``) and committing the
resulting input/output vectors under ``tests/golden/``; ``tests/test_oracle_golden.py``
checks the restatement against every one of them bit for bit (NMS) / to 0 ulp or
stated tolerance (decode).
"""
