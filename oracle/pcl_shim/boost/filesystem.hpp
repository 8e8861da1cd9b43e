// oracle/pcl_shim — main.cpp:7 includes boost/filesystem.hpp but uses nothing from it.
#pragma once
