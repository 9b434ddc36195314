// engine_solver.cu -- matrix-free implicit path (placeholder until the solver lands in this file).
#include "engine.hpp"
