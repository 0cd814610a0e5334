"""API-compatible mirrors of the reference ``models`` package members on the hot path."""
from .detector import LeakDetector  # noqa: F401
from .utils import WDNGraph, build_wdn_graph_from_inp, parse_epanet_inp  # noqa: F401
