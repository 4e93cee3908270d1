"""`_elf._options`: OptionSpec / OptionMap with the behaviour of the reference's C++ classes
(src_cpp/elf/options/OptionSpec.{h,cc}, OptionMap.{h,cc}) as seen through their pybind surface:
typed add*Option / add*ListOption with or without a default, the argparse description JSON
(`--name`, or `--no_name` with store_false for a bool that defaults to true), merge (existing names
win), prefix/suffix renaming; the map is a JSON object with load/get/set by JSON string and
"<name> has not been set!" for missing keys."""
import copy
import json

_PYTYPE = {"Int": "int", "Int32": "int", "UnsignedInt32": "int", "Int64": "int", "UnsignedInt64": "int",
           "Float": "float", "Float64": "float", "Float32": "float", "Bool": "bool", "Str": "str"}
_NO_DEFAULT = object()


class _Option:
    def __init__(self, name, help_, typename, is_list, default):
        self.name, self.help, self.typename, self.is_list, self.default = name, help_, typename, is_list, default

    def argparse(self):
        args = ["--" + self.name]
        has_default = self.default is not _NO_DEFAULT
        if self.typename == "bool" and not self.is_list:
            kwargs = {"help": self.help, "dest": self.name}
            flip = has_default and bool(self.default)
            kwargs["action"] = "store_false" if flip else "store_true"
            if flip:
                args = ["--no_" + self.name]
            return {"args": args, "kwargs": kwargs}
        kwargs = {"type": self.typename, "help": self.help, "required": not has_default, "dest": self.name}
        if has_default:
            kwargs["default"] = self.default
        if self.is_list:
            kwargs["nargs"] = "*"
        return {"args": args, "kwargs": kwargs}


class OptionSpec:
    def __init__(self, other=None):
        # std::unordered_map in the reference: iteration order is unspecified there, insertion order here
        self._opts = {} if other is None else {k: copy.copy(v) for k, v in other._opts.items()}

    def _add(self, typename, is_list, name, help_, default=_NO_DEFAULT):
        if name in self._opts:
            return False  # emplace().second
        if default is not _NO_DEFAULT:
            default = list(default) if is_list else default
        self._opts[name] = _Option(name, help_, typename, is_list, default)
        return True

    def getOptionNames(self):
        return list(self._opts)

    def getPythonArgparseOptionsAsJSONString(self):
        return json.dumps([o.argparse() for o in self._opts.values()])

    def merge(self, other):
        for k, v in other._opts.items():
            self._opts.setdefault(k, v)  # unordered_map::insert keeps existing keys

    def addPrefixSuffixToOptionNames(self, prefix, suffix):
        new = {}
        for k, v in self._opts.items():
            v.name = prefix + v.name + suffix
            new[prefix + k + suffix] = v
        self._opts = new


def _make_adders():
    for cname, tname in _PYTYPE.items():
        def scalar(self, name, help_, default=_NO_DEFAULT, _t=tname):
            return self._add(_t, False, name, help_, default)

        def lst(self, name, help_, default=_NO_DEFAULT, _t=tname):
            return self._add(_t, True, name, help_, default)
        setattr(OptionSpec, f"add{cname}Option", scalar)
        setattr(OptionSpec, f"add{cname}ListOption", lst)


_make_adders()


class OptionMap:
    def __init__(self, spec_or_map):
        if isinstance(spec_or_map, OptionMap):
            self._spec = OptionSpec(spec_or_map._spec)
            self._data = copy.deepcopy(spec_or_map._data)
        else:
            self._spec = OptionSpec(spec_or_map)
            self._data = {}

    def getOptionSpec(self):
        return self._spec

    def getJSONString(self):
        return json.dumps(self._data)

    def loadJSONString(self, s):
        self._data.update(json.loads(s))

    def getAsJSONString(self, name):
        if name not in self._data:
            raise RuntimeError(name + " has not been set!")
        return json.dumps(self._data[name])

    def setAsJSONString(self, name, s):
        self._data[name] = json.loads(s)
