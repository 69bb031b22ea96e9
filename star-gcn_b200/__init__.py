"""stargcn_b200 — Blackwell-native hot path of STAR-GCN.

The package directory is ``star-gcn_b200/``; it is importable as ``stargcn_b200`` through
the one-file shim ``stargcn_b200.py`` at the repo root.

    seg_op     the eight ``mx.nd.contrib.seg_*`` operators of the reference on torch tensors
    layers     mirror of ``mxgraph.layers`` (aggregators, HeterGCNLayer, StackedHeterGCNLayers)
    graph      device-resident multi-relation CSR plans + the fused aggregation op
    decoder    masked-embedding lookup, reconstruction decoder and losses
    sampler    device-resident graph: neighbour sampling, level split, support, edge removal, id merging
    devgraph   device-resident heterogeneous graph + gen_plan without host round trips of index data
    dist       node-partitioned multi-GPU aggregation (halo exchange over NCCL)
    runtime    CUDA-graph step capture and stream fork/join
    optim      multi-tensor global-norm clip + Adam
    static_step  a whole training iteration on static whole-graph plans as one CUDA graph
    model      the encoder-decoder stack of experiments/STAR-GCN.py:Net assembled from the pieces above
"""
from . import _lib  # noqa: F401  (fails loudly if the CUDA library is missing)
from . import seg_op, graph, layers, decoder, sampler, devgraph, runtime, optim, model, static_step  # noqa: F401

__version__ = "0.1.0"
