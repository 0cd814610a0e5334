"""ORACLE -- test infrastructure only.

CPU restatement of the reference's message-passing path.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs
may import this package; the product (``leak_det_gnn_b200``) never does.
"""
