// stub: the reference's gpu_objects.h only needs the ANARIDataType name
#pragma once
typedef int ANARIDataType;
