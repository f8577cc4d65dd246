"""Noise module: same public names as the reference's ``src/noise/__init__.py:9-28`` where they are
on the GPU path (``run_trial_fast``, ``CompiledCircuit``, ``build_decoding_matrices``)."""
from .builder import build_decoding_matrices, build_fault_tables, fault_tables_for
from .compiled import CompiledCircuit

__all__ = ["run_trial_fast", "CompiledCircuit", "build_decoding_matrices", "build_fault_tables", "fault_tables_for"]


def __getattr__(name):          # lazy: importing the package must not require the CUDA library
    if name in ("run_trial_fast", "run_trials_from_events", "events_from_random"):
        from . import simulation
        return getattr(simulation, name)
    if name in ("simulate_circuit_Z", "simulate_circuit_X", "sparsify_syndrome", "extract_data_qubit_state",
                "generate_noisy_circuit"):
        from . import twins
        return getattr(twins, name)
    raise AttributeError(name)
