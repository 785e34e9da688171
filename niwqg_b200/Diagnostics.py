"""Diagnostics registry: same protocol as the reference's niwqg/Diagnostics.py
(``add_diagnostic``, ``increment_diagnostics``, ``get_diagnostic``), host-side
bookkeeping only.  The numbers come from one C-ABI call per tick
(``niwqg_diagnostics``) issued by the model's ``_calc_derived_fields``.
"""
import numpy as np


def add_diagnostic(self, diag_name, description=None, units=None, types='scalar', function=None):
    """niwqg/Diagnostics.py:13-24."""
    assert hasattr(function, '__call__')
    assert isinstance(diag_name, str)
    self.diagnostics[diag_name] = {'description': description, 'units': units, 'active': True, 'count': 0,
                                   'type': types, 'function': function}


def get_diagnostic(self, dname):
    """niwqg/Diagnostics.py:6-8."""
    return self.diagnostics[dname]['value'] / self.diagnostics[dname]['count']


def describe_diagnostics(self):
    """niwqg/Diagnostics.py:26-35 (made Python-3 safe)."""
    print('NAME               | DESCRIPTION')
    print(80 * '-')
    for k in sorted(self.diagnostics.keys()):
        print('{:<10} | {:<54}'.format(k, self.diagnostics[k]['description']))


def increment_diagnostics(self):
    """niwqg/Diagnostics.py:41-58: every ``tdiags`` steps (tested on the step counter
    BEFORE it advances, so the 'time' series lags one step - F10) evaluate every
    registered function and append scalars."""
    if not (self.tc % self.tdiags):
        self._calc_derived_fields()
        for dname, d in self.diagnostics.items():
            res = d['function'](self)
            if d['type'] == 'scalar':
                if 'value' in d:
                    if np.ndim(res) == 0:
                        d['value'] = np.hstack([d['value'], res])
                    else:                       # ensemble: one row per tick
                        d['value'] = np.vstack([d['value'], np.asarray(res)[None, :]])
                else:
                    d['value'] = np.array(res) if np.ndim(res) == 0 else np.asarray(res)[None, :]
            else:
                if 'value' in d:
                    d['value'] += res
                    d['value'] *= 0.5
                else:
                    d['value'] = res
