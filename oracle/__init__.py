"""CPU oracle for the EMIP motion-stream hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``emip_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs use it, and only as the checker.

Parity pin: the reference (zhangxin06/EMIP) ships no tests or golden vectors
(SURVEY.md section 4).  The oracle is therefore pinned against outputs of the
reference's own Python code imported in the build container
(``tests/golden/make_golden.py`` -> ``tests/golden/*.pt``); the not-gpu test
suite re-checks every restatement here against those committed vectors.
"""
