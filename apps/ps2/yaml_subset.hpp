// The subset of YAML that config/ps2.yaml uses: a two-level block mapping of scalars
//     key: value
//     section:
//       key: value
// with `#` comments and the `---` / `...` document markers (config/ps2.yaml:1-41).  The reference
// parses the file with yaml-cpp (external/yaml-cpp, ProblemSets/ps2_cpp/lib/Config.cpp:36-50); this
// path needs none of the rest of YAML.
#pragma once
#include <fstream>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>

namespace ps2 {

class YamlDoc {
public:
    // "section.key" -> scalar text; top-level scalars are stored under their own name.
    std::map<std::string, std::string> values;

    static YamlDoc load(const std::string& path) {
        std::ifstream in(path);
        if (!in) throw std::runtime_error("could not open " + path);
        std::stringstream ss; ss << in.rdbuf();
        return parse(ss.str());
    }

    static YamlDoc parse(const std::string& text) {
        YamlDoc doc;
        std::istringstream in(text);
        std::string line, section;
        int lineno = 0;
        while (std::getline(in, line)) {
            ++lineno;
            const std::string body = strip_comment(line);
            const size_t first = body.find_first_not_of(" \t");
            if (first == std::string::npos) continue;
            const std::string t = trim(body);
            if (t == "---" || t == "...") continue;
            const size_t colon = find_key_colon(t);
            if (colon == std::string::npos) throw std::runtime_error("yaml line " + std::to_string(lineno) + ": expected `key: value`");
            const std::string key = trim(t.substr(0, colon)), val = unquote(trim(t.substr(colon + 1)));
            if (first == 0) {
                if (val.empty()) section = key;              // opens a nested mapping
                else { section.clear(); doc.values[key] = val; }
            } else {
                if (section.empty()) throw std::runtime_error("yaml line " + std::to_string(lineno) + ": indented entry outside a section");
                doc.values[section + "." + key] = val;
            }
        }
        return doc;
    }

    bool has(const std::string& key) const { return values.count(key) != 0; }
    bool has_section(const std::string& section) const {
        auto it = values.lower_bound(section + ".");
        return it != values.end() && it->first.compare(0, section.size() + 1, section + ".") == 0;
    }
    std::string str(const std::string& key) const {
        auto it = values.find(key);
        if (it == values.end()) throw std::runtime_error("missing config key " + key);
        return it->second;
    }
    long integer(const std::string& key) const {
        const std::string v = str(key);
        size_t pos = 0;
        long r = std::stol(v, &pos);
        if (pos != v.size()) throw std::runtime_error("config key " + key + ": not an integer: " + v);
        return r;
    }
    bool boolean(const std::string& key) const {              // yaml-cpp's bool spellings
        std::string v = str(key);
        for (auto& c : v) c = char(std::tolower(static_cast<unsigned char>(c)));
        if (v == "true" || v == "yes" || v == "on" || v == "y") return true;
        if (v == "false" || v == "no" || v == "off" || v == "n") return false;
        throw std::runtime_error("config key " + key + ": not a boolean: " + v);
    }

private:
    static std::string trim(const std::string& s) {
        const size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
        return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
    }
    static std::string strip_comment(const std::string& s) {
        bool sq = false, dq = false;
        for (size_t i = 0; i < s.size(); ++i) {
            if (s[i] == '\'' && !dq) sq = !sq;
            else if (s[i] == '"' && !sq) dq = !dq;
            else if (s[i] == '#' && !sq && !dq && (i == 0 || s[i - 1] == ' ' || s[i - 1] == '\t')) return s.substr(0, i);
        }
        return s;
    }
    static size_t find_key_colon(const std::string& t) {       // first ':' followed by space or end of line
        for (size_t i = 0; i < t.size(); ++i)
            if (t[i] == ':' && (i + 1 == t.size() || t[i + 1] == ' ' || t[i + 1] == '\t')) return i;
        return std::string::npos;
    }
    static std::string unquote(const std::string& v) {
        if (v.size() >= 2 && ((v.front() == '"' && v.back() == '"') || (v.front() == '\'' && v.back() == '\''))) return v.substr(1, v.size() - 2);
        return v;
    }
};

} // namespace ps2
