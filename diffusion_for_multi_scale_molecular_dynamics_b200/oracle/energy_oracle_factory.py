"""Energy oracle factory (reference: oracle/energy_oracle_factory.py:8-49): ``name: lammps`` selects the
Stillinger-Weber oracle, evaluated on the GPU here instead of through LAMMPS."""
from typing import Any, AnyStr, Dict, List

from diffusion_for_multi_scale_molecular_dynamics_b200.oracle.energy_oracle import EnergyOracle, OracleParameters
from diffusion_for_multi_scale_molecular_dynamics_b200.oracle.stillinger_weber_energy_oracle import (
    LammpsOracleParameters, StillingerWeberEnergyOracle)

ORACLE_PARAMETERS_BY_NAME = dict(lammps=LammpsOracleParameters)
ENERGY_ORACLE_BY_NAME = dict(lammps=StillingerWeberEnergyOracle)


def create_energy_oracle_parameters(energy_oracle_dictionary: Dict[AnyStr, Any], elements: List[str]) -> OracleParameters:
    """Configuration object for the oracle named in a yaml ``oracle`` block."""
    name = energy_oracle_dictionary["name"]
    assert name in ORACLE_PARAMETERS_BY_NAME.keys(), \
        f"Energy Oracle {name} is not implemented. Possible choices are {ORACLE_PARAMETERS_BY_NAME.keys()}"
    return ORACLE_PARAMETERS_BY_NAME[name](**energy_oracle_dictionary, elements=elements)


def create_energy_oracle(oracle_parameters: OracleParameters) -> EnergyOracle:
    """Instantiate the energy oracle."""
    name = oracle_parameters.name
    assert name in ENERGY_ORACLE_BY_NAME.keys(), \
        f"Energy Oracle {name} is not implemented. Possible choices are {ENERGY_ORACLE_BY_NAME.keys()}"
    return ENERGY_ORACLE_BY_NAME[name](oracle_parameters)
