#pragma once
class CommandQueue;
