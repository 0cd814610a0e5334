"""leak_det_gnn_b200 -- B200-native (sm_100a) message-passing path for the Leak-det-gnn detector.

Public surface (mirrors what the reference's hot path touches):

* ``leak_det_gnn_b200.models.detector.LeakDetector``   drop-in for reference models/detector.py
* ``leak_det_gnn_b200.models.utils.build_wdn_graph_from_inp``  drop-in for models/utils.py:84-166
* ``leak_det_gnn_b200.nn.GCNConv`` / ``global_mean_pool``      drop-ins for the two PyG operators
* ``leak_det_gnn_b200.lib``                                    ctypes binding of the C ABI (include/ltgnn.h)
"""
__version__ = "0.1.0"
