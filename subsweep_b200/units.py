"""SI values of the units and constants the sweep path uses (src/units/mod.rs:16-108).

All quantities cross the C ABI as f64 in SI base units, like the reference's diman quantities.
"""
METERS = 1.0
CENTIMETERS = 0.01
PARSEC = 3.0857e16
KILOPARSEC = 1000 * PARSEC
MEGAPARSEC = 1000000 * PARSEC
SECONDS = 1.0
YEARS = 3.15576e7
KILOYEARS = 1000.0 * YEARS
MEGAYEARS = 1e6 * YEARS
KILOGRAMS = 1.0
GRAMS = 1e-3
CUBIC_CENTIMETERS = 1e-6
GRAMS_PER_CUBIC_CENTIMETER = GRAMS / CUBIC_CENTIMETERS
PER_CUBIC_CENTIMETER = 1.0 / CUBIC_CENTIMETERS
PROTON_MASS = 1.67262192369e-27
BOLTZMANN_CONSTANT = 1.380649e-23

_UNITS = {
    "": 1.0, "s": SECONDS, "yr": YEARS, "kyr": KILOYEARS, "Myr": MEGAYEARS, "Gyr": 1e9 * YEARS,
    "m": METERS, "cm": CENTIMETERS, "pc": PARSEC, "kpc": KILOPARSEC, "Mpc": MEGAPARSEC,
    "s^-1": 1.0, "K": 1.0, "%": 0.01,
}


def parse_quantity(text) -> float:
    """'1 Myr', '1.0e-5 s^-1', 0.1 -> SI f64 (the subset of diman's parser the sweep section needs)."""
    if isinstance(text, (int, float)):
        return float(text)
    parts = str(text).split()
    value = float(parts[0])
    unit = parts[1] if len(parts) > 1 else ""
    if unit not in _UNITS:
        raise ValueError(f"unknown unit {unit!r} in {text!r}")
    return value * _UNITS[unit]
