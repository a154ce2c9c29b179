// main_clumpy.cc -- CLI dispatch with the behaviour of the reference's main (main_clumpy.cc:13-60
// there): `help` / no argument lists the commands, `examples` prints one usage example per command,
// anything else is looked up in the registry and run; exit status 0 on success, 1 otherwise.
#include "clumpy_command.hh"

#include <cstdio>
#include <cstring>
#include <memory>

static const char* kBanner =
    "\nWelcome to clumpy. This tool can process or generate large swaths of image\n"
    "data. It is meant to be used in conjunction with Python.\n";

static void print_help() {
    printf("%s\n", kBanner);
    printf("%-20s %s\n", "help", "list all commands and their arguments");
    printf("%-20s %s\n", "examples", "print a usage example for each command");
    for (const auto& entry : ClumpyCommand::registry()) {
        std::unique_ptr<ClumpyCommand> cmd(entry.second());
        printf("%-20s %s\n", entry.first.c_str(), cmd->description().c_str());
        if (!cmd->usage().empty()) printf("%-20s usage: %s %s\n", "", entry.first.c_str(), cmd->usage().c_str());
    }
}

static void print_examples() {
    for (const auto& entry : ClumpyCommand::registry()) {
        std::unique_ptr<ClumpyCommand> cmd(entry.second());
        printf("clumpy %s %s\n", entry.first.c_str(), cmd->example().c_str());
    }
}

int main(int argc, const char* argv[]) {
    if (argc <= 1 || !strcmp(argv[1], "help")) { print_help(); return 0; }
    if (!strcmp(argv[1], "examples")) { print_examples(); return 0; }
    auto& reg = ClumpyCommand::registry();
    auto found = reg.find(argv[1]);
    if (found == reg.end()) { print_help(); return 1; }
    std::unique_ptr<ClumpyCommand> cmd(found->second());
    std::vector<std::string> args(argv + 2, argv + argc);
    return cmd->exec(args) ? 0 : 1;
}
