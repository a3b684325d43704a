// TEST INFRASTRUCTURE ONLY - empty stand-in: include/Auxiliar.h includes this g2o header but the frontend path uses nothing of it.
#pragma once
