"""Same surface as the reference's src/network/__init__.py."""
from .network import DensityNetwork


def get_network(type):
    if type == "mlp":
        return DensityNetwork
    else:
        raise NotImplementedError("Unknown network type!")
