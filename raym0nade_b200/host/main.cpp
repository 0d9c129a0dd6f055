// main.cpp — the console program (main.cpp of the reference): one command per line until `exit` (or end of input,
// where the reference would spin).
#include <iostream>
#include <string>

#include "host.hpp"

int main() {
    std::cout << "raym0nade on " << rm_version() << std::endl;
    MyConsole console;
    std::string opt;
    while (true) {
        std::cout << "> ";
        if (!std::getline(std::cin, opt)) break;
        if (!opt.empty() && opt.back() == '\r') opt.pop_back();
        if (opt == "exit") break;
        if (opt.empty()) continue;
        parseCommand(console, opt);
    }
    std::cout << std::endl;
    return 0;
}
