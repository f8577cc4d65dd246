"""``generate_noisy_circuit`` on tuple circuits (reference ``src/noise/model.py:4-58``); a host-side twin kept for code
that imports it -- shots on the GPU path draw faults in kernel K1 and never build a noisy circuit."""
from .twins import generate_noisy_circuit

__all__ = ["generate_noisy_circuit"]
