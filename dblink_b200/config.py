"""HOCON subset parser for dblink configuration files (Run.scala:35 uses typesafe-config).

Covers what the reference's configs use (examples/RLdata500.conf, docs/configuration.md): `//` and `#` comments,
unquoted or quoted keys, `:` or `=` (or nothing before `{`) as separators, nested objects, arrays, newline or comma
as element separator, trailing commas, `${a.b.c}` substitutions (resolved against the root after parsing, like
ConfigFactory.resolve()), dotted keys (`a.b : 1`), strings, numbers, booleans, null.  Not supported: include,
`+=`, string concatenation of several unquoted tokens with substitutions, triple-quoted strings.
"""
import re


class ConfigError(ValueError):
    pass


class _Subst:
    def __init__(self, path, optional):
        self.path, self.optional = path, optional


_NUM = re.compile(r"^-?(\d+\.?\d*([eE][-+]?\d+)?|\.\d+([eE][-+]?\d+)?)$")


class _Parser:
    def __init__(self, text):
        self.s, self.i, self.n = text, 0, len(text)

    def error(self, msg):
        line = self.s.count("\n", 0, self.i) + 1
        raise ConfigError(f"line {line}: {msg}")

    def ws(self, newlines=True):
        while self.i < self.n:
            c = self.s[self.i]
            if c in " \t\r" or (newlines and c == "\n"):
                self.i += 1
            elif c == "#" or self.s.startswith("//", self.i):
                while self.i < self.n and self.s[self.i] != "\n":
                    self.i += 1
            else:
                break

    def parse_root(self):
        self.ws()
        if self.i < self.n and self.s[self.i] == "{":
            v = self.parse_object()
        else:
            v = self.parse_members(end=None)
        self.ws()
        if self.i != self.n:
            self.error("trailing characters")
        return v

    def parse_object(self):
        assert self.s[self.i] == "{"
        self.i += 1
        v = self.parse_members(end="}")
        return v

    def parse_members(self, end):
        obj = {}
        while True:
            self.ws()
            if self.i >= self.n:
                if end is None:
                    return obj
                self.error("unterminated object")
            c = self.s[self.i]
            if end is not None and c == end:
                self.i += 1
                return obj
            if c == ",":
                self.i += 1
                continue
            key = self.parse_key()
            self.ws(newlines=False)
            if self.i < self.n and self.s[self.i] in ":=":
                self.i += 1
                self.ws()
                val = self.parse_value()
            elif self.i < self.n and self.s[self.i] == "{":
                val = self.parse_object()
            else:
                self.error(f"expected ':' '=' or '{{' after key {key!r}")
            self._set(obj, key, val)

    @staticmethod
    def _set(obj, key, val):
        parts = key if isinstance(key, list) else [key]
        for p in parts[:-1]:
            nxt = obj.get(p)
            if not isinstance(nxt, dict):
                nxt = {}
                obj[p] = nxt
            obj = nxt
        last = parts[-1]
        if isinstance(val, dict) and isinstance(obj.get(last), dict):
            _merge(obj[last], val)  # HOCON: duplicate object keys merge
        else:
            obj[last] = val

    def parse_key(self):
        if self.s[self.i] == '"':
            return [self.parse_string()]
        j = self.i
        while j < self.n and self.s[j] not in ":={}[],\n \t\r\"#" and not self.s.startswith("//", j):
            j += 1
        if j == self.i:
            self.error("empty key")
        key = self.s[self.i:j]
        self.i = j
        return key.split(".")

    def parse_string(self):
        assert self.s[self.i] == '"'
        self.i += 1
        out = []
        while self.i < self.n and self.s[self.i] != '"':
            c = self.s[self.i]
            if c == "\\":
                self.i += 1
                esc = self.s[self.i]
                out.append({"n": "\n", "t": "\t", "r": "\r", '"': '"', "\\": "\\", "/": "/"}.get(esc, esc))
            else:
                out.append(c)
            self.i += 1
        if self.i >= self.n:
            self.error("unterminated string")
        self.i += 1
        return "".join(out)

    def parse_value(self):
        c = self.s[self.i] if self.i < self.n else ""
        if c == "{":
            return self.parse_object()
        if c == "[":
            return self.parse_array()
        if c == '"':
            return self.parse_string()
        if self.s.startswith("${", self.i):
            j = self.s.find("}", self.i)
            if j < 0:
                self.error("unterminated substitution")
            body = self.s[self.i + 2:j]
            self.i = j + 1
            opt = body.startswith("?")
            return _Subst(body[1:] if opt else body, opt)
        j = self.i
        while j < self.n and self.s[j] not in ",}]\n#" and not self.s.startswith("//", j):
            j += 1
        tok = self.s[self.i:j].strip()
        self.i = j
        if tok == "":
            self.error("missing value")
        if tok == "true":
            return True
        if tok == "false":
            return False
        if tok == "null":
            return None
        if _NUM.match(tok):
            return int(tok) if re.match(r"^-?\d+$", tok) else float(tok)
        return tok  # unquoted string

    def parse_array(self):
        assert self.s[self.i] == "["
        self.i += 1
        out = []
        while True:
            self.ws()
            if self.i >= self.n:
                self.error("unterminated array")
            c = self.s[self.i]
            if c == "]":
                self.i += 1
                return out
            if c == ",":
                self.i += 1
                continue
            out.append(self.parse_value())


def _merge(a, b):
    for k, v in b.items():
        if isinstance(v, dict) and isinstance(a.get(k), dict):
            _merge(a[k], v)
        else:
            a[k] = v


def _lookup(root, path):
    cur = root
    for p in path.split("."):
        if not isinstance(cur, dict) or p not in cur:
            raise KeyError(path)
        cur = cur[p]
    return cur


def _resolve(node, root, stack=()):
    if isinstance(node, _Subst):
        if node.path in stack:
            raise ConfigError(f"substitution cycle at ${{{node.path}}}")
        try:
            target = _lookup(root, node.path)
        except KeyError:
            if node.optional:
                return None
            raise ConfigError(f"could not resolve substitution ${{{node.path}}}") from None
        import copy

        return _resolve(copy.deepcopy(target), root, stack + (node.path,))
    if isinstance(node, dict):
        for k in list(node):
            node[k] = _resolve(node[k], root, stack)
        return node
    if isinstance(node, list):
        return [_resolve(v, root, stack) for v in node]
    return node


class Config:
    """Minimal typesafe-config-like accessor: getString/getInt/... by dotted path."""

    def __init__(self, tree):
        self.tree = tree

    def has(self, path):
        try:
            _lookup(self.tree, path)
            return True
        except KeyError:
            return False

    def get(self, path, default=KeyError):
        try:
            return _lookup(self.tree, path)
        except KeyError:
            if default is KeyError:
                raise ConfigError(f"No configuration setting found for key '{path}'") from None
            return default

    def get_config(self, path):
        v = self.get(path)
        if not isinstance(v, dict):
            raise ConfigError(f"{path} is not an object")
        return Config(v)

    def get_string(self, path):
        return str(self.get(path))

    def get_int(self, path):
        return int(self.get(path))

    def get_double(self, path):
        return float(self.get(path))

    def get_bool(self, path):
        v = self.get(path)
        if isinstance(v, bool):
            return v
        raise ConfigError(f"{path} is not a boolean")

    def get_list(self, path):
        v = self.get(path)
        if not isinstance(v, list):
            raise ConfigError(f"{path} is not a list")
        return v


def parse_string(text):
    tree = _Parser(text).parse_root()
    return Config(_resolve(tree, tree))


def parse_file(path):
    with open(path) as fh:
        return parse_string(fh.read())
