"""dsmil_wsi_b200 -- B200-native DSMIL aggregator hot path.

Public surface == the reference's dsmil.py: FCLayer, IClassifier, BClassifier, MILNet.
"""
from .modules import BClassifier, FCLayer, IClassifier, MILNet  # noqa: F401

__all__ = ["FCLayer", "IClassifier", "BClassifier", "MILNet"]
