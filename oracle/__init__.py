"""CPU oracle for the DTQN hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A CPU restatement (numpy / plain torch fp32) of the reference algorithm (kevslinger/DTQN) for the path
named by BASELINE.json: CarFlag / Memory envs, TimeLimit, replay buffer, acting context, DTQN forward,
double-DQN TD step, grad clip and Adam.  Every function cites the reference file:line it follows.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may
import this package, and there only as the checker or the reported CPU baseline.  The product
(``dtqn_b200``) never imports it and fails loudly when its CUDA library is missing.

Parity pinning: the reference has no tests or golden vectors, so the oracle is pinned by
EXECUTING the reference in the build container (``oracle/ref_harness``) -- ``tests/golden/gen_golden.py``
writes the fixtures under ``tests/golden/`` and ``tests/test_oracle_vs_golden.py`` checks this port
against them (and against the live reference when /root/reference is present).
"""
