"""Mirror of the graph part of reference models/utils.py (lines 18-166)."""
from ..graph import WDNGraph, build_wdn_graph_from_inp, parse_epanet_inp  # noqa: F401

__all__ = ["WDNGraph", "build_wdn_graph_from_inp", "parse_epanet_inp"]
