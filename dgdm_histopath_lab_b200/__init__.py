"""dgdm_histopath_lab_b200 — B200-native (sm_100a) implementation of the DGDM forward hot path of
danieleschmidt/dgdm-histopath-lab behind the reference's Python API.

    from dgdm_histopath_lab_b200 import DGDMModel, TissueGraphBuilder, Data, Batch

The package directory is named with underscores (a hyphen is not importable); it corresponds to the
`dgdm-histopath-lab_b200/` package of the build brief.
"""
from .errors import ModelConfigurationError, ModelInferenceError, ValidationError
from .graph import Batch, Data
from .core import (AdaptiveGraphPooling, DiffusionLayer, DiffusionScheduler, DynamicGraphLayer, FeatureEncoder,
                   GraphConvolution, GraphEncoder, GraphUNet, MultiHeadAttention, SpatialAttention)
from .model import (ClassificationHead, DGDMModel, GlobalAttentionPool, GlobalMaxPool, GlobalMeanPool, GlobalSet2SetPool,
                    RegressionHead)
from .builder import GraphEdge, GraphNode, PatchInfo, SlideData, TissueGraphBuilder
from .losses import ContrastiveLoss
from .io import GraphFileDataset, collate_to_device, load_graph, save_graph

__all__ = [
    "DGDMModel", "TissueGraphBuilder", "Data", "Batch", "DiffusionLayer", "DiffusionScheduler", "GraphConvolution",
    "DynamicGraphLayer", "AdaptiveGraphPooling", "GraphUNet", "MultiHeadAttention", "SpatialAttention", "FeatureEncoder",
    "GraphEncoder", "ClassificationHead", "RegressionHead", "GlobalAttentionPool", "GlobalMeanPool", "GlobalMaxPool",
    "GlobalSet2SetPool", "ModelConfigurationError", "ModelInferenceError", "ValidationError", "PatchInfo", "SlideData",
    "GraphNode", "GraphEdge", "ContrastiveLoss", "GraphFileDataset", "collate_to_device", "load_graph", "save_graph",
]
__version__ = "0.1.0"
