// console.cpp — the interactive front end: named models and named render arguments, created, inspected, deleted and
// rendered by one-line commands.  The surface is the reference's (include/myconsole.h; the dialogue of
// docs/renderArguments.txt drives this binary unchanged: same commands, same prompt order, same replies), the
// construction is this host's own: commands live in a dispatch table, the two kinds of named object share one set of
// reply builders, and the render arguments are read through a table of field descriptors from whatever stream the
// console was given - so a script, a pipe or a test can hold the other end.
#include <array>
#include <iostream>
#include <limits>
#include <sstream>

#include "host.hpp"

namespace {

// ---- replies: "<Kind> (<name>) <what>."  The two kinds differ only in how the reference happens to spell them.
struct Kind { const char *stored, *missing; };
constexpr Kind kModel{"Model", "Model"}, kArgs{"RenderArgs", "Args"};

void say(std::ostream &out, const char *kind, const std::string &name, const char *what) {
    out << kind << " (" << name << ") " << what << std::endl;
}
void sayExists(std::ostream &out, const Kind &k, const std::string &name) { say(out, k.missing, name, "is already exists."); }
void sayMissing(std::ostream &out, const Kind &k, const std::string &name) { say(out, k.missing, name, "does not exists."); }

// ---- render-argument fields in the order the dialogue asks for them (src/myconsole.cpp:27-67)
struct Prompted {
    const char *prompt;
    bool (*read)(std::istream &, RenderArgs &);
};
bool readVec(std::istream &in, vec3 &v) { return bool(in >> v.x >> v.y >> v.z); }
const std::array<Prompted, 8> kArgFields{{
    {"direction (x,y,z): ", [](std::istream &in, RenderArgs &a) { return readVec(in, a.direction); }},
    {"right (x,y,z): ", [](std::istream &in, RenderArgs &a) { return readVec(in, a.right); }},
    {"up (x,y,z): ", [](std::istream &in, RenderArgs &a) { return readVec(in, a.up); }},
    // the camera position is entered as coefficients of the three axes just read
    {"position (D,R,U): ", [](std::istream &in, RenderArgs &a) {
         vec3 c;
         if (!readVec(in, c)) return false;
         a.position = c.x * a.direction + c.y * a.right + c.z * a.up;
         return true;
     }},
    {"accuracy, focus, CoC, exposure: ", [](std::istream &in, RenderArgs &a) { return bool(in >> a.accuracy >> a.focus >> a.CoC >> a.exposure); }},
    {"width, height: ", [](std::istream &in, RenderArgs &a) { return bool(in >> a.width >> a.height); }},
    {"spp, threads, P_Direct: : ", [](std::istream &in, RenderArgs &a) { return bool(in >> a.spp >> a.threads >> a.P_Direct); }},
    {"savePath: ", [](std::istream &in, RenderArgs &a) { return bool(in >> a.savePath); }},
}};

void restOfLine(std::istream &in) { in.ignore(std::numeric_limits<std::streamsize>::max(), '\n'); }

void show(std::ostream &out, const char *label, const vec3 &v) { out << label << " : " << v.x << " " << v.y << " " << v.z << std::endl; }

// ---- commands: "<verb> <noun> <name>" or "render <model> <args>"
struct Command {
    const char *verb, *noun;                           // noun == nullptr: the verb takes two names instead
    void (*run)(MyConsole &, const std::string &, const std::string &);
};
const Command kCommands[] = {
    {"create", "model", [](MyConsole &c, const std::string &n, const std::string &) { c.createModel(n); }},
    {"create", "args", [](MyConsole &c, const std::string &n, const std::string &) { c.createRenderArgs(n); }},
    {"delete", "model", [](MyConsole &c, const std::string &n, const std::string &) { c.deleteModel(n); }},
    {"delete", "args", [](MyConsole &c, const std::string &n, const std::string &) { c.deleteRenderArgs(n); }},
    {"view", "model", [](MyConsole &c, const std::string &n, const std::string &) { c.viewModel(n); }},
    {"view", "args", [](MyConsole &c, const std::string &n, const std::string &) { c.viewRenderArgs(n); }},
    {"render", nullptr, [](MyConsole &c, const std::string &m, const std::string &a) { c.render(m, a); }},
};

}  // namespace

MyConsole::MyConsole() : MyConsole(std::cin, std::cout) {}
MyConsole::MyConsole(std::istream &in, std::ostream &out) : in_(in), out_(out) {}

void MyConsole::createModel(const std::string &model_id) {
    if (models.count(model_id)) return sayExists(out_, kModel, model_id);
    std::array<std::string, 3> answer;                 // folder, file, sky map
    const char *ask[3] = {"Enter the model path (e.g., fbx/): ", "Enter the model name (e.g., model.rmscene): ", "Enter the sky map name (e.g., sky.hdr): "};
    for (int k = 0; k < 3; k++) {
        out_ << ask[k];
        in_ >> answer[k];
    }
    if (!in_) return say(out_, kModel.stored, model_id, "was not created: input ended.");
    restOfLine(in_);
    // built inside the map node: a Model must never move once loaded (it is pointed into, like the reference's)
    models.emplace(std::piecewise_construct, std::forward_as_tuple(model_id), std::forward_as_tuple(answer[0], answer[1], answer[2]));
    say(out_, kModel.stored, model_id, "created.");
}

void MyConsole::createRenderArgs(const std::string &name) {
    if (renderArgs.count(name)) return sayExists(out_, kArgs, name);
    RenderArgs fresh;
    bool complete = true;
    for (const Prompted &f : kArgFields) {
        out_ << f.prompt;
        if (complete && !f.read(in_, fresh)) complete = false;
    }
    if (!complete) {                                   // a half-read record is dropped rather than rendered from
        in_.clear();
        restOfLine(in_);
        return say(out_, kArgs.stored, name, "was not created: could not read all 22 fields.");
    }
    restOfLine(in_);
    renderArgs.emplace(name, fresh);
    say(out_, kArgs.stored, name, "created.");
}

void MyConsole::deleteModel(const std::string &name) {
    if (!models.erase(name)) return sayMissing(out_, kModel, name);
    say(out_, kModel.stored, name, "deleted.");
}

void MyConsole::deleteRenderArgs(const std::string &name) {
    if (!renderArgs.erase(name)) return sayMissing(out_, kArgs, name);
    say(out_, kArgs.stored, name, "deleted.");
}

void MyConsole::viewModel(const std::string &name) {
    const auto it = models.find(name);
    if (it == models.end()) return sayMissing(out_, kModel, name);
    out_ << "Model Path: " << it->second.model_path << std::endl << "Faces: " << it->second.faceCount() << std::endl;
}

void MyConsole::viewRenderArgs(const std::string &name) {
    const auto it = renderArgs.find(name);
    if (it == renderArgs.end()) return sayMissing(out_, kArgs, name);
    const RenderArgs &a = it->second;
    show(out_, "direction", a.direction);
    show(out_, "right", a.right);
    show(out_, "up", a.up);
    show(out_, "position", a.position);
    out_ << "accuracy: " << a.accuracy << std::endl << "exposure: " << a.exposure << std::endl
         << "width, height: " << a.width << " " << a.height << std::endl << "spp: " << a.spp << std::endl
         << "threads: " << a.threads << std::endl << "P_Direct: " << a.P_Direct << std::endl << "savePath: " << a.savePath << std::endl;
}

void MyConsole::render(const std::string &model_name, const std::string &args_name) {
    const auto m = models.find(model_name);
    if (m == models.end()) return sayMissing(out_, kModel, model_name);
    const auto a = renderArgs.find(args_name);
    if (a == renderArgs.end()) return sayMissing(out_, kArgs, args_name);
    render_multiThread(m->second, a->second);
}

void parseCommand(MyConsole &console, const std::string &line) {
    std::istringstream words(line);
    std::string verb, second, third;
    words >> verb >> second >> third;
    bool verbKnown = false;
    for (const Command &c : kCommands) {
        if (verb != c.verb) continue;
        verbKnown = true;
        if (!c.noun) return c.run(console, second, third);
        if (second == c.noun) return c.run(console, third, std::string());
    }
    if (!verbKnown) console.out() << "Unknown command." << std::endl;      // a known verb with an unknown noun is silently ignored, as in the reference
}

int runConsole(MyConsole &console, std::istream &in, std::ostream &out) {
    for (std::string line; out << "> ", std::getline(in, line);) {
        if (!line.empty() && line.back() == '\r') line.pop_back();
        if (line == "exit") break;
        if (!line.empty()) parseCommand(console, line);
    }
    out << std::endl;
    return 0;
}
