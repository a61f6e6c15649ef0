"""Minimal stand-in for the `prettytable` package (absent from this image, no network) covering what the
reference's benchmark scripts use (/root/reference/tools/benchmark/pt_bench.py:329-399, ncu_bench.py:255-463):
`PrettyTable(field_names=...)`, `.field_names`, `.align[col]`, `.add_row`, `str(table)`, `.get_csv_string`.
Only put on sys.path when the real package is missing."""


class PrettyTable:
    def __init__(self, field_names=None):
        self.field_names = list(field_names) if field_names else []
        self.align = {}
        self.rows = []

    def add_row(self, row):
        self.rows.append([str(x) for x in row])

    def get_csv_string(self, header=True):
        lines = [",".join(self.field_names)] if header else []
        lines += [",".join(r) for r in self.rows]
        return "\r\n".join(lines) + "\r\n"

    def get_string(self):
        cols = [self.field_names] + self.rows
        widths = [max(len(r[i]) for r in cols) for i in range(len(self.field_names))]

        def fmt(row):
            cells = []
            for i, c in enumerate(row):
                a = self.align.get(self.field_names[i], "c")
                cells.append(c.ljust(widths[i]) if a == "l" else c.rjust(widths[i]) if a == "r" else c.center(widths[i]))
            return "| " + " | ".join(cells) + " |"

        bar = "+" + "+".join("-" * (w + 2) for w in widths) + "+"
        return "\n".join([bar, fmt(self.field_names), bar] + [fmt(r) for r in self.rows] + [bar])

    __str__ = get_string
