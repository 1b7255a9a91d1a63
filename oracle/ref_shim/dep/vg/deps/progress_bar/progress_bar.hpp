// TEST INFRASTRUCTURE ONLY -- shim: silent progress bar (src/rmdup.cpp:81 creates one).
#pragma once
#include <string>
class ProgressBar {
public:
    ProgressBar(unsigned long, const char*) {}
    void Progressed(unsigned long) {}
};
