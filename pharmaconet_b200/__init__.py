"""pharmaconet_b200 - B200-native implementation of PharmacoNet's virtual-screening hot path."""

__version__ = "0.1.0"
