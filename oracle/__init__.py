"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the caption-generation hot path.

Nothing in the product package (``video-captioning_b200/``) may import this package.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs use it, and only as the checker / reported CPU baseline.

Contents
--------
``caption_oracle.py``  closed-form CPU restatement (plain matmul + sigmoid/tanh, no nn.LSTM) of
                       the reference path; every function cites the reference file:line it follows.
``synth.py``           deterministic (numpy PCG64) synthetic state_dicts / features shared by the
                       oracle, the golden-vector generator, the tests and bench.py.
``ref_shim.py``        loads the UNMODIFIED reference sources from /root/reference behind an import
                       shim.  Only usable where /root/reference exists
                       (the build container); never on the GPU box.
``make_golden.py``     regenerates tests/golden/*.npz from the real reference via ref_shim.

Parity pinning: the reference has no tests or golden vectors of its own, so
the oracle is pinned against *outputs of the reference itself run in the build container*
(``make_golden.py`` -> ``tests/golden``), see DESIGN.md "Oracle".
"""
