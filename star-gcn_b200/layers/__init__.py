"""Mirror of ``mxgraph.layers`` (mxgraph/layers/__init__.py:1-3)."""
from .common import *  # noqa: F401,F403
from .common import Dense, activation_code, get_activation
from .aggregators import BaseAggregator, GCNAggregator, MultiLinkGCNAggregator
from .layers import HeterGCNLayer, InnerProductLayer, LayerDictionary, StackedHeterGCNLayers
