// main.cpp — the console program: commands from standard input until `exit` or end of input.
#include <iostream>

#include "host.hpp"

int main() {
    std::cout << "raym0nade on " << rm_version() << std::endl;
    MyConsole console(std::cin, std::cout);
    return runConsole(console, std::cin, std::cout);
}
