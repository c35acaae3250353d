"""Layers complying with connectivity restrictions (mirror of reference cpflow/topology.py:7-20, 36-38)."""


def connected_layer(num_qubits):
    """All-to-all pairs (reference topology.py:7-8)."""
    return [[i, j] for i in range(num_qubits) for j in range(i + 1, num_qubits)]


def chain_layer(num_qubits):
    """Nearest-neighbour chain (reference topology.py:11-12)."""
    return [[i, i + 1] for i in range(num_qubits - 1)]


def fill_layers(layer, depth):
    """`depth` blocks: complete repetitions of `layer` plus a remainder (reference topology.py:15-20)."""
    num_complete_layers = depth // len(layer)
    complete_layers = [layer, num_complete_layers]
    incomplete_layer = layer[:depth % len(layer)]
    return {'layers': complete_layers, 'free': incomplete_layer}


def num_qubits_from_layer(layer):
    """Largest qubit index in the coupling map plus one (reference topology.py:36-38)."""
    return max([item for sublist in layer for item in sublist]) + 1
