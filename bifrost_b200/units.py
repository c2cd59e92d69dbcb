"""Tiny unit converter for the quantities the hot-path blocks touch
(frequency, time, dispersion measure).  The reference uses `pint` (python/bifrost/units.py:28),
which is not a dependency here."""
_FACTORS = {
    'Hz': 1.0, 'kHz': 1e3, 'MHz': 1e6, 'GHz': 1e9, 'THz': 1e12,
    's': 1.0, 'ms': 1e-3, 'us': 1e-6, 'ns': 1e-9,
    '1/s': 1.0, 's^-1': 1.0,
}
_KIND = {'Hz': 'f', 'kHz': 'f', 'MHz': 'f', 'GHz': 'f', 'THz': 'f', '1/s': 'f', 's^-1': 'f',
         's': 't', 'ms': 't', 'us': 't', 'ns': 't'}
# dispersion measure (blocks/fdmt.py and the sigproc sink): the spellings pint accepts
for _dm in ('pc cm^-3', 'pc cm**-3', 'pc/cm^3', 'pc/cm**3', 'pc / cm ** 3', 'parsec / centimeter ** 3'):
    _FACTORS[_dm], _KIND[_dm] = 1.0, 'dm'
_INV = {'s': 'Hz', 'ms': 'kHz', 'us': 'MHz', 'ns': 'GHz', 'Hz': 's', 'kHz': 'ms', 'MHz': 'us', 'GHz': 'ns'}


def convert_units(value, old_units, new_units):
    if old_units == new_units:
        return value
    if old_units not in _FACTORS or new_units not in _FACTORS or _KIND[old_units] != _KIND[new_units]:
        raise ValueError(f"Cannot convert '{old_units}' to '{new_units}'")
    return value * _FACTORS[old_units] / _FACTORS[new_units]


def transform_units(units, exponent):
    """units ** exponent for exponent = -1 (what blocks/fft.py needs)."""
    if exponent == 1 or units is None:
        return units
    if exponent == -1 and units in _INV:
        return _INV[units]
    return f"({units})^{exponent}"
