#pragma once
namespace accelerated { struct Future { void wait() const {} }; }
