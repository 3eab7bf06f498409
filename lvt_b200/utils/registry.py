"""Name -> class registry with the fvcore `Registry` surface the reference uses
(vidgen/utils/registry.py:2; `@REG.register()` and `REG.get(name)`)."""


class Registry:
    def __init__(self, name):
        self._name = name
        self._obj_map = {}

    def _do_register(self, name, obj):
        if name in self._obj_map:
            raise KeyError(f"An object named '{name}' was already registered in '{self._name}' registry!")
        self._obj_map[name] = obj

    def register(self, obj=None):
        if obj is None:
            def deco(fn_or_cls):
                self._do_register(fn_or_cls.__name__, fn_or_cls)
                return fn_or_cls
            return deco
        self._do_register(obj.__name__, obj)
        return obj

    def get(self, name):
        if name not in self._obj_map:
            raise KeyError(f"No object named '{name}' found in '{self._name}' registry!")
        return self._obj_map[name]

    def __contains__(self, name):
        return name in self._obj_map
