"""Component / Port plumbing of the drop-in classes.

When ISCE2 is importable the real ``iscesys.Component.Component`` is used (so ``configure()``, catalog recording
and pickling behave exactly as in the host application).  Otherwise a small stand-in provides the part of that
interface the zero-Doppler components and their callers use (components/iscesys/Component/Component.py:162-360,
Configurable.py:213-310): ``Parameter`` declarations, ``configure()``, ``wireInputPort(name=, object=)``,
``_inputPorts.getPort(name).getObject()``, ``activateInputPorts()``, ``createPorts()``, ``logger``.
"""
from __future__ import annotations

import logging

try:  # pragma: no cover - exercised only inside an ISCE2 installation
    from iscesys.Component.Component import Component, Port  # type: ignore

    HAVE_ISCE = True
except Exception:  # ISCE2 is not installed: stand-in
    HAVE_ISCE = False

    class Port:
        """components/iscesys/Component/Component.py:162-200"""

        def __init__(self, name=None, method=None, doc=None):
            self._name = name
            self._method = method
            self._object = None
            self.__doc__ = doc

        def getName(self):
            return self._name

        def getMethod(self):
            return self._method

        def setObject(self, obj):
            self._object = obj

        def getObject(self):
            return self._object

        def __call__(self, *args, **kwargs):
            return self._method(*args, **kwargs)

        name = property(getName)
        object = property(getObject, setObject)

    class PortIterator:
        """components/iscesys/Component/Component.py:203-280"""

        def __init__(self):
            self._ports = {}

        def add(self, port):
            self._ports[port.getName()] = port

        def getPort(self, name=None):
            try:
                return self._ports[name]
            except KeyError:
                raise KeyError(f"No port named {name} found")

        def hasPort(self, name=None):
            return name in self._ports

        def __iter__(self):
            return iter(self._ports.values())

        def __setitem__(self, name, method):  # old-style: self.inputPorts['frame'] = self.addFrame
            self.add(Port(name=name, method=method))

        def __getitem__(self, name):
            return self._ports[name].getObject()

        def __contains__(self, name):
            return name in self._ports

    class Parameter:
        """components/iscesys/Component/Configurable.py:1077 (Configurable.Parameter)"""

        def __init__(self, attrname, public_name="", default=None, container=None, type=type, mandatory=False,
                     units=None, doc="", private=False, intent="input"):
            self.attrname = attrname
            self.public_name = public_name
            self.default = default
            self.container = container
            self.type = type
            self.mandatory = mandatory
            self.units = units
            self.doc = doc
            self.private = private
            self.intent = intent

    class Component:
        family = "component"
        logging_name = "isce.component"
        parameter_list = ()
        Parameter = Parameter

        def __init__(self, family=None, name=None):
            self.family = family or self.__class__.family
            self.name = name or ""
            self._inputPorts = PortIterator()
            self._outputPorts = PortIterator()
            self.logger = logging.getLogger(self.logging_name)
            self.dictionaryOfVariables = {}
            self.dictionaryOfOutputVariables = {}
            self.descriptionOfVariables = {}
            self.mandatoryVariables = []
            self.optionalVariables = []
            for par in self.parameter_list:
                setattr(self, par.attrname, par.default)
            self.createPorts()

        # --- Configurable surface -------------------------------------------------------
        def configure(self):
            for par in self.parameter_list:
                if not hasattr(self, par.attrname):
                    setattr(self, par.attrname, par.default)
            return self

        def initOptionalAndMandatoryLists(self):
            for key, val in self.dictionaryOfVariables.items():
                if isinstance(val, (list, tuple)) and len(val) > 2:
                    (self.mandatoryVariables if val[2] == "mandatory" else self.optionalVariables).append(key)

        # --- ports ----------------------------------------------------------------------
        @property
        def inputPorts(self):
            return self._inputPorts

        def createPorts(self):
            pass

        def wireInputPort(self, name=None, object=None):
            """components/iscesys/Component/Component.py:348-360"""
            if not self._inputPorts.hasPort(name):
                raise KeyError(f"No input port named {name} found")
            self._inputPorts.getPort(name).setObject(object)

        def activateInputPorts(self):
            for port in self._inputPorts:
                port()

        def listInputPorts(self):
            return [p.getName() for p in self._inputPorts]
