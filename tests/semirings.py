"""The 960 built-in semirings of the reference (Source/axb.m:19-45, Source/axb_template.m:25-86,
Source/axb_compare_template.m:4-66; count derivation Source/GB_AxB_Gustavson_builtin.c:147-200), by
their public names GxB_<ADD>_<MULT>_<TYPE> (Include/GraphBLAS.h:5238-5497)."""

NONBOOL = ["INT8", "UINT8", "INT16", "UINT16", "INT32", "UINT32", "INT64", "UINT64", "FP32", "FP64"]
TT_OPS = ["FIRST", "SECOND", "MIN", "MAX", "PLUS", "MINUS", "TIMES", "DIV", "ISEQ", "ISNE", "ISGT",
          "ISLT", "ISGE", "ISLE", "LOR", "LAND", "LXOR"]
CMP_OPS = ["EQ", "NE", "GT", "LT", "GE", "LE"]
NUM_MONOIDS = ["MIN", "MAX", "PLUS", "TIMES"]
BOOL_MONOIDS = ["LOR", "LAND", "LXOR", "EQ"]
BOOL_OPS = ["FIRST", "SECOND", "LOR", "LAND", "LXOR", "EQ", "GT", "LT", "GE", "LE"]


def all_builtin():
    """-> list of (add, mult, xytype); 680 + 240 + 40 = 960 unique workers"""
    out = []
    for t in NONBOOL:
        for add in NUM_MONOIDS:
            for m in TT_OPS:
                out.append((add, m, t))
        for add in BOOL_MONOIDS:
            for m in CMP_OPS:
                out.append((add, m, t))
    for add in BOOL_MONOIDS:
        for m in BOOL_OPS:
            out.append((add, m, "BOOL"))
    return out


def name(add, mult, t):
    return f"GxB_{add}_{mult}_{t}"
