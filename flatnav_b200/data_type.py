"""`flatnav.data_type` of the reference (python-bindings/src/flatnav/bindings.cpp:507-515).

Values are those of flatnav::util::DataType (include/flatnav/util/Datatype.h:11-24); they are part of
the index file format.
"""
import enum


class DataType(enum.IntEnum):
    uint8 = 0
    int8 = 4
    float32 = 9


# py::enum_::export_values() also puts the members on the module
uint8 = DataType.uint8
int8 = DataType.int8
float32 = DataType.float32
