#pragma once
namespace slam { struct MapPointRecord {}; }
