"""nuwa_pytorch_b200 -- B200-native (sm_100a) implementation of the NUWA hot paths.

Public surface mirrors nuwa_pytorch/__init__.py:1-5 of the reference for the in-scope classes.
"""
__version__ = "0.1.0"
