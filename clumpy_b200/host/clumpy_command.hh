// clumpy_command.hh -- the command plugin interface of clumpy, kept API-identical so that command
// translation units written against the reference (clumpy_command.hh:8-29 there) compile unchanged:
// an abstract ClumpyCommand with usage/description/example/exec, a name -> factory registry held in
// a function-local static (safe under static-initialisation order), and a Register helper whose
// constructor adds a factory; each command TU defines one static Register object.
#pragma once

#include <functional>
#include <string>
#include <unordered_map>
#include <vector>

struct ClumpyCommand {
    using FactoryFn = std::function<ClumpyCommand*()>;
    using Registry = std::unordered_map<std::string, FactoryFn>;

    virtual ~ClumpyCommand() {}

    // one-line argument synopsis, one-line description, example argument list
    virtual std::string usage() const = 0;
    virtual std::string description() const = 0;
    virtual std::string example() const = 0;

    // argv[2..] by value; false makes the process exit with status 1
    virtual bool exec(std::vector<std::string> args) = 0;

    static Registry& registry() {
        static Registry commands;
        return commands;
    }

    struct Register {
        Register(std::string id, FactoryFn make) { registry()[id] = make; }
    };
};
