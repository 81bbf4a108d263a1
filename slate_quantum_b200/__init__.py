"""slate_quantum_b200: B200-native (sm_100a) implementation of slate_quantum's Bloch build -> eigh -> evolve path."""

__version__ = "0.1.0"
