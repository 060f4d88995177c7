"""TEST INFRASTRUCTURE ONLY — CPU oracle for the quantized-MoE decode hot path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  The product package
(``ktransformers_b200``) never does, and fails loudly when its CUDA library is missing.

Contents
--------
``ktoracle.c/.h``   plain-C restatement of the reference arithmetic (ggml block formats, Q8_K /
                    Q8_0 activation quantisation, integer dot products, MOE/Linear/MLP forward).
``bindings.py``     ctypes faces of ``libktoracle.so`` (ours) and ``_ref/libktref_<isa>.so``
                    (the unmodified reference compiled by ``oracle/Makefile``).
``gate_oracle.py``  numpy restatement of DeepSeek ``MoEGate.forward`` routing.
``mla_oracle.py``   numpy/torch fp32 restatement of absorbed-MLA paged decode attention.

Parity status: PINNED against ``oracle/_ref`` and the fixtures in ``tests/golden``.
"""
