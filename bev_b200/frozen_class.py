"""Attribute-freezing base class (mirrors /root/reference/bev/frozen_class.py:1-10).

After ``_freeze()`` assigning an attribute that does not already exist raises ``TypeError``;
existing attributes stay writable.  ``Calib`` and ``BEVWorldSpec`` derive from it.
"""


class FrozenClass(object):
    _frozen = False

    def __setattr__(self, name, value):
        if self._frozen and not hasattr(self, name):
            raise TypeError("%r is a frozen class" % self)
        object.__setattr__(self, name, value)

    def _freeze(self):
        object.__setattr__(self, "_frozen", True)
