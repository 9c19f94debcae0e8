// Minimal XML reader for URDF / SRDF text: elements, attributes, comments, declarations.  No external
// dependency (urdfdom / tinyxml2 are not part of this image).  Text content is ignored -- URDF and SRDF carry
// everything in attributes.
#pragma once
#include <map>
#include <memory>
#include <string>
#include <vector>

namespace b200sv
{
struct XmlNode
{
  std::string name;
  std::map<std::string, std::string> attrs;
  std::vector<std::unique_ptr<XmlNode>> children;

  const XmlNode* child(const std::string& n) const
  {
    for (const auto& c : children)
      if (c->name == n)
        return c.get();
    return nullptr;
  }
  std::vector<const XmlNode*> all(const std::string& n) const
  {
    std::vector<const XmlNode*> out;
    for (const auto& c : children)
      if (c->name == n)
        out.push_back(c.get());
    return out;
  }
  bool has(const std::string& a) const { return attrs.count(a) != 0; }
  std::string attr(const std::string& a, const std::string& def = "") const
  {
    auto it = attrs.find(a);
    return it == attrs.end() ? def : it->second;
  }
};

class XmlParser
{
public:
  // Returns the root element or nullptr (err set).
  static std::unique_ptr<XmlNode> parse(const std::string& text, std::string& err)
  {
    XmlParser p(text);
    std::unique_ptr<XmlNode> root;
    while (p.skipMisc())
    {
      if (p.peek() != '<')
      {
        ++p.pos_;  // stray text
        continue;
      }
      root = p.element();
      break;
    }
    if (!root || !p.err_.empty())
    {
      err = p.err_.empty() ? "no root element" : p.err_;
      return nullptr;
    }
    return root;
  }

private:
  explicit XmlParser(const std::string& t) : s_(t) {}
  const std::string& s_;
  size_t pos_ = 0;
  std::string err_;

  char peek() const { return pos_ < s_.size() ? s_[pos_] : '\0'; }
  bool starts(const char* lit) const { return s_.compare(pos_, std::char_traits<char>::length(lit), lit) == 0; }
  void skipWs()
  {
    while (pos_ < s_.size() && (s_[pos_] == ' ' || s_[pos_] == '\t' || s_[pos_] == '\n' || s_[pos_] == '\r'))
      ++pos_;
  }
  // skip whitespace, comments, <? ?> and <! > ; returns false at end of input
  bool skipMisc()
  {
    for (;;)
    {
      skipWs();
      if (pos_ >= s_.size())
        return false;
      if (starts("<!--"))
      {
        size_t e = s_.find("-->", pos_ + 4);
        pos_ = e == std::string::npos ? s_.size() : e + 3;
      }
      else if (starts("<?"))
      {
        size_t e = s_.find("?>", pos_ + 2);
        pos_ = e == std::string::npos ? s_.size() : e + 2;
      }
      else if (starts("<!"))
      {
        size_t e = s_.find('>', pos_ + 2);
        pos_ = e == std::string::npos ? s_.size() : e + 1;
      }
      else
        return true;
    }
  }
  static bool nameChar(char c)
  {
    return (c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z') || (c >= '0' && c <= '9') || c == '_' || c == '-' ||
           c == ':' || c == '.';
  }
  std::string name()
  {
    size_t b = pos_;
    while (pos_ < s_.size() && nameChar(s_[pos_]))
      ++pos_;
    return s_.substr(b, pos_ - b);
  }
  static std::string unescape(const std::string& v)
  {
    if (v.find('&') == std::string::npos)
      return v;
    std::string o;
    for (size_t i = 0; i < v.size(); ++i)
    {
      if (v[i] != '&')
      {
        o += v[i];
        continue;
      }
      static const char* ent[] = { "&amp;", "&lt;", "&gt;", "&quot;", "&apos;" };
      static const char rep[] = { '&', '<', '>', '"', '\'' };
      bool hit = false;
      for (int k = 0; k < 5 && !hit; ++k)
      {
        size_t n = std::char_traits<char>::length(ent[k]);
        if (v.compare(i, n, ent[k]) == 0)
        {
          o += rep[k];
          i += n - 1;
          hit = true;
        }
      }
      if (!hit)
        o += v[i];
    }
    return o;
  }
  std::unique_ptr<XmlNode> element()
  {
    if (peek() != '<')
    {
      err_ = "expected '<'";
      return nullptr;
    }
    ++pos_;
    auto node = std::make_unique<XmlNode>();
    node->name = name();
    if (node->name.empty())
    {
      err_ = "empty element name at offset " + std::to_string(pos_);
      return nullptr;
    }
    for (;;)
    {
      skipWs();
      if (starts("/>"))
      {
        pos_ += 2;
        return node;
      }
      if (peek() == '>')
      {
        ++pos_;
        break;
      }
      std::string an = name();
      if (an.empty())
      {
        err_ = "malformed attribute in <" + node->name + ">";
        return nullptr;
      }
      skipWs();
      if (peek() != '=')
      {
        err_ = "attribute without value in <" + node->name + ">";
        return nullptr;
      }
      ++pos_;
      skipWs();
      char q = peek();
      if (q != '"' && q != '\'')
      {
        err_ = "unquoted attribute in <" + node->name + ">";
        return nullptr;
      }
      size_t e = s_.find(q, pos_ + 1);
      if (e == std::string::npos)
      {
        err_ = "unterminated attribute in <" + node->name + ">";
        return nullptr;
      }
      node->attrs[an] = unescape(s_.substr(pos_ + 1, e - pos_ - 1));
      pos_ = e + 1;
    }
    // children until </name>
    for (;;)
    {
      // skip text
      while (pos_ < s_.size() && s_[pos_] != '<')
        ++pos_;
      if (!skipMisc())
      {
        err_ = "unexpected end inside <" + node->name + ">";
        return nullptr;
      }
      if (peek() != '<')
        continue;
      if (starts("</"))
      {
        pos_ += 2;
        std::string cn = name();
        if (cn != node->name)
        {
          err_ = "mismatched </" + cn + "> for <" + node->name + ">";
          return nullptr;
        }
        skipWs();
        if (peek() == '>')
          ++pos_;
        return node;
      }
      auto c = element();
      if (!c)
        return nullptr;
      node->children.push_back(std::move(c));
    }
  }
};
}  // namespace b200sv
