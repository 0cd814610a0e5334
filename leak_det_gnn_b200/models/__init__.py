"""API-compatible mirrors of the reference ``models`` package members on (and next to) the hot path."""
from .detector import LeakDetector  # noqa: F401
from .predictor import NormalPredictorGRU, NormalPredictorTCN  # noqa: F401
from .utils import (WDNGraph, build_residual_sequence_from_segment, build_wdn_graph_from_inp,  # noqa: F401
                    parse_epanet_inp)
