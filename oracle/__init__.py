"""TEST INFRASTRUCTURE ONLY — CPU oracle for the STAR-GCN aggregation hot path.

Nothing under ``oracle/`` is part of the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it, and there only as the checker (or the CPU arm being timed), never
as a fallback for the CUDA path.

Modules
-------
segops   ctypes front-end of ``seg_ops_oracle.c`` (plain-C restatement of
         /root/reference/seg_ops_cuda/mxnet_op/seg_op.cc:7-332)
ref      ctypes front-end of ``oracle/_ref/*.so`` — the reference's OWN CPU code compiled
         from /root/reference (seg_ops.cu CPU loops, GraphSampler bookkeeping)
npy_ref  the numpy known-answer functions of the reference's test_seg_ops.py, loaded
         from /root/reference at fixture-generation time only
layers   numpy restatement of MultiLinkGCNAggregator / HeterGCNLayer / decoder maths
graphs   synthetic bipartite rating graphs of the BASELINE.json shapes

Parity pin status: PINNED for the segment operators (``_ref`` + ``npy_ref`` fixtures in
tests/golden/); "parity unpinned" for the MXNet-resident pieces (FullyConnected, LeakyReLU,
Embedding, L2 loss) whose arithmetic lives in un-vendored Apache MXNet 1.5.x — see
DESIGN.md §Oracle.
"""
