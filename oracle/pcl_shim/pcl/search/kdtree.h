#pragma once
#include <pcl/search/search.h>
