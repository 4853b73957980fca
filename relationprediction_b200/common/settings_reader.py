"""Experiment-settings (.exp) reader with the reference's file format and lookup semantics
(common/settings_reader.py:29-48 in the reference): tab-indented nesting, `[Section]` headers,
`key=value` lines, every value kept as a string, blank lines ignored.  Written from the format
description; tests compare its parse of the shipped .exp files with the reference reader's."""


class Settings(object):
    def __init__(self, values=None):
        self._v = dict(values or {})

    # dict-like surface the rest of the reference code relies on
    def __getitem__(self, key):
        return self._v[key]

    def __contains__(self, key):
        return key in self._v

    def __iter__(self):
        return iter(self._v)

    def __repr__(self):
        return repr(self._v)

    __str__ = __repr__

    def put(self, key, value):
        self._v[key] = value

    def merge(self, other):
        """Copy (overwrite) every top-level entry of `other` into this section (train.py:80-86)."""
        for k in other:
            self._v[k] = other[k]

    def to_dict(self):
        return {k: (v.to_dict() if isinstance(v, Settings) else v) for k, v in self._v.items()}


def _depth(line):
    n = 0
    while n < len(line) and line[n] == "\t":
        n += 1
    return n


def parse_lines(lines):
    root = Settings()
    stack = [(0, root)]  # (depth of the entries this section holds, section)
    for raw in lines:
        if not raw.strip():
            continue
        depth = _depth(raw)
        text = raw.strip()
        while len(stack) > 1 and depth < stack[-1][0]:
            stack.pop()
        if depth > stack[-1][0]:
            # deeper than the open section expects: the reference skips such lines
            continue
        section = stack[-1][1]
        if text.startswith("["):
            child = Settings()
            section.put(text[1:-1], child)
            stack.append((depth + 1, child))
        else:
            parts = [p.strip() for p in text.split("=")]
            section.put(parts[0], parts[1])
    return root


def read(filename):
    with open(filename, "r") as fh:
        return parse_lines(list(fh))


def read_string(text):
    return parse_lines(text.splitlines(True))
