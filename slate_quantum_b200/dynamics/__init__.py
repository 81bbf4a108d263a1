"""Mirror of slate_quantum.dynamics for the eigenbasis path (reference: slate_quantum/dynamics/__init__.py)."""

from slate_quantum_b200.dynamics.schrodinger import (
    solve_schrodinger_equation,
    solve_schrodinger_equation_decomposition,
)

__all__ = ["solve_schrodinger_equation", "solve_schrodinger_equation_decomposition"]
