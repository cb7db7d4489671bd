// Text helpers for the flame compiler: the macro grammar of variations.yaml.
// Behavioural restatement of the reference's regex helpers (src/util.cpp:6-23,
// src/variation_table.cpp:9-16) without std::regex.
#pragma once
#include <set>
#include <string>
#include <vector>

namespace rfk {

// `$name` followed by a non-identifier character becomes `value` + that character.
// The trailing character is consumed by the match (so a `$name` at the very end of
// the text is NOT replaced, and the scan resumes after the consumed character),
// exactly as std::regex_replace does with "\\$name([^a-zA-Z0-9_])" (util.cpp:6-9).
std::string replace_macro(const std::string& str, const std::string& name, const std::string& value);

// All `$[a-z0-9_]+` names in the text (util.cpp:11-23).
std::set<std::string> find_macros(const std::string& str);

// Plain substring replacement, left to right, non-overlapping (variation_table.cpp:9-16).
std::string replace_all(std::string str, const std::string& from, const std::string& to);

// Whole file as a string; ok=false when the file cannot be opened.
std::string read_file(const std::string& path, bool* ok = nullptr);

// Whitespace-separated tokens (std::istringstream >> semantics, util.hpp:12-25).
std::vector<std::string> split_ws(const std::string& s);

}  // namespace rfk
