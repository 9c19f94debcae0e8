// URDF + SRDF -> flat robot / collision model, built the way MoveIt 2 builds them.  Reference (moveit_core/):
//   robot_model/src/robot_model.cpp:848-890    buildRecursive (DFS pre-order; one joint + one link per URDF link)
//   robot_model/src/robot_model.cpp:895-1044   jointBoundsFromURDF / constructJointModel (root joint from SRDF)
//   robot_model/src/robot_model.cpp:1165-1254  urdfPose2Isometry3d / constructLinkModel
//   robot_model/src/robot_model.cpp:404-471    buildMimic
//   robot_model/src/robot_model.cpp:136-169    computeDescendantsHelper (follows mimic requests)
//   robot_model/src/robot_model.cpp:538-598, 728-845  buildGroups / addJointModelGroup
//   robot_model/src/joint_model_group.cpp:60-92, 114-300  includesParent / JointModelGroup ctor
//   collision_detection/src/collision_matrix.cpp:67-78   ACM from SRDF
//   collision_distance_field/src/collision_env_distance_field.cpp:735-838, 907-953  DFCE masks, self-field points
//   collision_distance_field/src/collision_distance_field_types.cpp:59-85, 292-335  BodyDecomposition (SPHERE shapes)
//   distance_field/src/find_internal_points.cpp:39-64     interior lattice points
// urdfdom behaviour relied upon (un-vendored): children of a link are ordered by joint name (std::map iteration
// in ModelInterface::initTree), Rotation::setFromRPY, axis default (1,0,0), mimic multiplier/offset defaults 1/0.
#include "host_model.hpp"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <functional>
#include <limits>
#include <map>
#include <set>
#include <sstream>

#include "xml_mini.hpp"

namespace b200sv
{
namespace
{
inline double& M(Mat4& m, int r, int c) { return m[c * 4 + r]; }
inline double Mc(const Mat4& m, int r, int c) { return m[c * 4 + r]; }

Mat4 identity()
{
  Mat4 m{};
  m[0] = m[5] = m[10] = m[15] = 1.0;
  return m;
}

void quatToRot(double x, double y, double z, double w, Mat4& m)  // Eigen::QuaternionBase::toRotationMatrix
{
  const double tx = 2.0 * x, ty = 2.0 * y, tz = 2.0 * z;
  const double twx = tx * w, twy = ty * w, twz = tz * w;
  const double txx = tx * x, txy = ty * x, txz = tz * x;
  const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
  M(m, 0, 0) = 1.0 - (tyy + tzz); M(m, 0, 1) = txy - twz; M(m, 0, 2) = txz + twy;
  M(m, 1, 0) = txy + twz; M(m, 1, 1) = 1.0 - (txx + tzz); M(m, 1, 2) = tyz - twx;
  M(m, 2, 0) = txz - twy; M(m, 2, 1) = tyz + twx; M(m, 2, 2) = 1.0 - (txx + tyy);
}

Mat4 poseToMatrix(const double* xyz, const double* rpy)  // urdfdom setFromRPY + normalize, then Translation * q
{
  const double phi = rpy[0] / 2.0, the = rpy[1] / 2.0, psi = rpy[2] / 2.0;
  double x = sin(phi) * cos(the) * cos(psi) - cos(phi) * sin(the) * sin(psi);
  double y = cos(phi) * sin(the) * cos(psi) + sin(phi) * cos(the) * sin(psi);
  double z = cos(phi) * cos(the) * sin(psi) - sin(phi) * sin(the) * cos(psi);
  double w = cos(phi) * cos(the) * cos(psi) + sin(phi) * sin(the) * sin(psi);
  const double s = sqrt(x * x + y * y + z * z + w * w);
  if (s == 0.0)
  {
    x = y = z = 0.0;
    w = 1.0;
  }
  else
  {
    x /= s; y /= s; z /= s; w /= s;
  }
  Mat4 m = identity();
  quatToRot(x, y, z, w, m);
  M(m, 0, 3) = xyz[0]; M(m, 1, 3) = xyz[1]; M(m, 2, 3) = xyz[2];
  return m;
}

bool parseDoubles(const std::string& s, int n, double* out)
{
  std::istringstream is(s);
  for (int i = 0; i < n; ++i)
    if (!(is >> out[i]))
      return false;
  return true;
}

void jointTransform(int type, const double* axis, const double* v, Mat4& d)
{
  d = identity();
  switch (type)
  {
    case REVOLUTE:  // revolute_joint_model.cpp:250-287 (x2_.. cached at :73-82)
    {
      const double x2 = axis[0] * axis[0], y2 = axis[1] * axis[1], z2 = axis[2] * axis[2];
      const double xy = axis[0] * axis[1], xz = axis[0] * axis[2], yz = axis[1] * axis[2];
      const double c = cos(v[0]), s = sin(v[0]);
      const double t = 1.0 - c;
      const double txy = t * xy, txz = t * xz, tyz = t * yz;
      const double zs = axis[2] * s, ys = axis[1] * s, xs = axis[0] * s;
      d[0] = t * x2 + c; d[1] = txy + zs; d[2] = txz - ys;
      d[4] = txy - zs; d[5] = t * y2 + c; d[6] = tyz + xs;
      d[8] = txz + ys; d[9] = tyz - xs; d[10] = t * z2 + c;
      break;
    }
    case PRISMATIC:  // prismatic_joint_model.cpp:124-146
      d[12] = axis[0] * v[0]; d[13] = axis[1] * v[0]; d[14] = axis[2] * v[0];
      break;
    case PLANAR:  // planar_joint_model.cpp:345-349
    {
      const double c = cos(v[2]), s = sin(v[2]);
      M(d, 0, 0) = c; M(d, 0, 1) = -s; M(d, 1, 0) = s; M(d, 1, 1) = c;
      M(d, 2, 2) = (1.0 - c) + c;
      M(d, 0, 3) = v[0]; M(d, 1, 3) = v[1];
      break;
    }
    case FLOATING:  // floating_joint_model.cpp:231-236
    {
      const double n = sqrt(v[3] * v[3] + v[4] * v[4] + v[5] * v[5] + v[6] * v[6]);
      quatToRot(v[3] / n, v[4] / n, v[5] / n, v[6] / n, d);
      M(d, 0, 3) = v[0]; M(d, 1, 3) = v[1]; M(d, 2, 3) = v[2];
      break;
    }
    default: break;
  }
}
}  // namespace

Mat4 mulAffine(const Mat4& a, const Mat4& b)
{
  Mat4 t{};
  for (int c = 0; c < 4; ++c)
  {
    for (int r = 0; r < 3; ++r)
    {
      double s = Mc(a, r, 0) * Mc(b, 0, c);
      s += Mc(a, r, 1) * Mc(b, 1, c);
      s += Mc(a, r, 2) * Mc(b, 2, c);
      s += Mc(a, r, 3) * Mc(b, 3, c);
      M(t, r, c) = s;
    }
    M(t, 3, c) = c == 3 ? 1.0 : 0.0;
  }
  return t;
}

bool isIdentityTransform(const double* m)  // link_model.cpp:64-66
{
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c)
      if (std::abs(m[c * 4 + r] - (r == c ? 1.0 : 0.0)) > 1e-12)
        return false;
  const double n = sqrt(m[12] * m[12] + m[13] * m[13] + m[14] * m[14]);
  return n < std::numeric_limits<double>::epsilon();
}

void hostForwardKinematics(const HostRobot& r, const std::vector<double>& pos, std::vector<Mat4>& out)
{
  out.assign(r.n_links, identity());
  for (int l = 0; l < r.n_links; ++l)
  {
    Mat4 origin;
    std::copy(r.joint_origin.begin() + 16 * l, r.joint_origin.begin() + 16 * (l + 1), origin.begin());
    const int type = r.joint_type[l];
    const bool origin_id = isIdentityTransform(origin.data());
    Mat4 jt;
    if (type != FIXED)
      jointTransform(type, &r.joint_axis[3 * l], &pos[r.first_var[l]], jt);
    if (r.parent[l] >= 0)
    {
      const Mat4& P = out[r.parent[l]];
      if (type == FIXED)
        out[l] = mulAffine(P, origin);
      else if (origin_id)
        out[l] = mulAffine(P, jt);
      else
        out[l] = mulAffine(mulAffine(P, origin), jt);
    }
    else
    {
      if (type == FIXED)
        out[l] = identity();
      else if (origin_id)
        out[l] = jt;
      else
        out[l] = mulAffine(origin, jt);
    }
  }
}

// -------------------------------------------------------------------------------------------------------------
namespace
{
struct UJoint
{
  std::string name, type, parent, child, mimic_joint;
  double xyz[3] = { 0, 0, 0 }, rpy[3] = { 0, 0, 0 }, axis[3] = { 1, 0, 0 };
  bool has_limit = false, has_safety = false, has_mimic = false;
  double lower = 0, upper = 0, soft_lower = 0, soft_upper = 0, mimic_mult = 1, mimic_off = 0;
};
struct UShape
{
  std::string kind;
  double radius = 0;
  double dims[3] = { 0, 0, 0 };  // cylinder: radius, length; box: x, y, z
  Mat4 origin;
};
struct Joint
{
  std::string name;
  int type = FIXED, index = 0, first_var = 0, n_vars = 0, parent_link = -1, child_link = -1;
  double axis[3] = { 1, 0, 0 };
  double lo = 0, hi = 0;
  int mimic = -1;
  double mimic_factor = 1, mimic_offset = 0;
  std::vector<int> mimic_requests;
  std::vector<int> descendant_links;
};
struct Link
{
  std::string name;
  int index = 0, parent = -1;
  std::vector<UShape> shapes;
  Mat4 origin;
  std::vector<int> child_joints;
};
struct Group
{
  std::string name;
  std::vector<int> joints;  // sorted joint indices
};
}  // namespace

bool loadUrdfSrdf(const std::string& urdf, const std::string& srdf, const std::string& group_name, double resolution,
                  HostRobot& robot, HostCollisionModel& cm, std::vector<double>& self_points, SelfLocalPoints& self_local,
                  SelfLocalPoints& link_points, std::string& err,
                  bool& parse_error)
{
  parse_error = true;
  std::string xerr;
  auto uroot = XmlParser::parse(urdf, xerr);
  if (!uroot || uroot->name != "robot")
  {
    err = "URDF: " + (uroot ? std::string("root element is not <robot>") : xerr);
    return false;
  }
  std::unique_ptr<XmlNode> sroot;
  if (!srdf.empty())
  {
    sroot = XmlParser::parse(srdf, xerr);
    if (!sroot || sroot->name != "robot")
    {
      err = "SRDF: " + (sroot ? std::string("root element is not <robot>") : xerr);
      return false;
    }
  }
  // ---- URDF elements ------------------------------------------------------------------------------------
  std::map<std::string, std::vector<UShape>> ulinks;
  std::vector<std::string> ulink_order;
  for (const XmlNode* le : uroot->all("link"))
  {
    std::vector<UShape> shapes;
    for (const XmlNode* ce : le->all("collision"))
    {
      const XmlNode* ge = ce->child("geometry");
      if (!ge || ge->children.empty())
        continue;
      double xyz[3] = { 0, 0, 0 }, rpy[3] = { 0, 0, 0 };
      if (const XmlNode* oe = ce->child("origin"))
      {
        if (oe->has("xyz") && !parseDoubles(oe->attr("xyz"), 3, xyz))
        {
          err = "URDF: bad collision origin xyz in link " + le->attr("name");
          return false;
        }
        if (oe->has("rpy") && !parseDoubles(oe->attr("rpy"), 3, rpy))
        {
          err = "URDF: bad collision origin rpy in link " + le->attr("name");
          return false;
        }
      }
      UShape s;
      s.kind = ge->children[0]->name;
      s.origin = poseToMatrix(xyz, rpy);
      if (s.kind == "sphere")
        s.radius = atof(ge->children[0]->attr("radius", "0").c_str());
      else if (s.kind == "cylinder")
      {
        s.dims[0] = atof(ge->children[0]->attr("radius", "0").c_str());
        s.dims[1] = atof(ge->children[0]->attr("length", "0").c_str());
      }
      else if (s.kind == "box" && !parseDoubles(ge->children[0]->attr("size"), 3, s.dims))
      {
        err = "URDF: bad box size in link " + le->attr("name");
        return false;
      }
      shapes.push_back(s);
    }
    ulinks[le->attr("name")] = shapes;
    ulink_order.push_back(le->attr("name"));
  }
  std::map<std::string, UJoint> ujoints;  // std::map: iteration in joint-name order, like urdfdom's initTree
  for (const XmlNode* je : uroot->all("joint"))
  {
    UJoint j;
    j.name = je->attr("name");
    j.type = je->attr("type");
    const XmlNode *pe = je->child("parent"), *ce = je->child("child");
    if (!pe || !ce)
    {
      err = "URDF: joint " + j.name + " lacks parent/child";
      return false;
    }
    j.parent = pe->attr("link");
    j.child = ce->attr("link");
    if (const XmlNode* oe = je->child("origin"))
    {
      if (oe->has("xyz") && !parseDoubles(oe->attr("xyz"), 3, j.xyz))
      {
        err = "URDF: bad origin xyz in joint " + j.name;
        return false;
      }
      if (oe->has("rpy") && !parseDoubles(oe->attr("rpy"), 3, j.rpy))
      {
        err = "URDF: bad origin rpy in joint " + j.name;
        return false;
      }
    }
    if (const XmlNode* ae = je->child("axis"))
      if (ae->has("xyz") && !parseDoubles(ae->attr("xyz"), 3, j.axis))
      {
        err = "URDF: bad axis in joint " + j.name;
        return false;
      }
    if (const XmlNode* le = je->child("limit"))
    {
      j.has_limit = true;
      j.lower = atof(le->attr("lower", "0").c_str());
      j.upper = atof(le->attr("upper", "0").c_str());
    }
    if (const XmlNode* se = je->child("safety_controller"))
    {
      j.has_safety = true;
      j.soft_lower = atof(se->attr("soft_lower_limit", "0").c_str());
      j.soft_upper = atof(se->attr("soft_upper_limit", "0").c_str());
    }
    if (const XmlNode* me = je->child("mimic"))
    {
      j.has_mimic = true;
      j.mimic_joint = me->attr("joint");
      j.mimic_mult = atof(me->attr("multiplier", "1").c_str());
      j.mimic_off = atof(me->attr("offset", "0").c_str());
    }
    if (!ulinks.count(j.parent) || !ulinks.count(j.child))
    {
      err = "URDF: joint " + j.name + " references an unknown link";
      return false;
    }
    ujoints[j.name] = j;
  }
  std::map<std::string, std::vector<const UJoint*>> children;
  std::set<std::string> is_child;
  for (const auto& kv : ujoints)
  {
    children[kv.second.parent].push_back(&kv.second);
    is_child.insert(kv.second.child);
  }
  std::vector<std::string> roots;
  for (const auto& n : ulink_order)
    if (!is_child.count(n))
      roots.push_back(n);
  if (roots.size() != 1)
  {
    err = "URDF: expected exactly one root link, found " + std::to_string(roots.size());
    return false;
  }
  parse_error = false;

  // ---- SRDF elements --------------------------------------------------------------------------------------
  struct VJoint { std::string name, child, frame, type; };
  std::vector<VJoint> vjoints;
  struct GroupCfg { std::string name; std::vector<std::pair<std::string, std::string>> chains; std::vector<std::string> joints, links, subgroups; };
  std::vector<GroupCfg> gcfgs;
  std::vector<std::pair<std::string, std::string>> disabled, enabled;
  if (sroot)
  {
    for (const XmlNode* v : sroot->all("virtual_joint"))
      vjoints.push_back({ v->attr("name"), v->attr("child_link"), v->attr("parent_frame"), v->attr("type") });
    for (const XmlNode* g : sroot->all("group"))
    {
      GroupCfg c;
      c.name = g->attr("name");
      for (const XmlNode* x : g->all("chain"))
        c.chains.push_back({ x->attr("base_link"), x->attr("tip_link") });
      for (const XmlNode* x : g->all("joint"))
        c.joints.push_back(x->attr("name"));
      for (const XmlNode* x : g->all("link"))
        c.links.push_back(x->attr("name"));
      for (const XmlNode* x : g->all("group"))
        c.subgroups.push_back(x->attr("name"));
      gcfgs.push_back(c);
    }
    for (const XmlNode* d : sroot->all("disable_collisions"))
      disabled.push_back({ d->attr("link1"), d->attr("link2") });
    for (const XmlNode* d : sroot->all("enable_collisions"))
      enabled.push_back({ d->attr("link1"), d->attr("link2") });
  }

  // ---- buildRecursive ---------------------------------------------------------------------------------------
  std::vector<Joint> joints;
  std::vector<Link> links;
  std::map<std::string, int> joint_index, link_index;
  std::function<int(int, const std::string&, const UJoint*)> build = [&](int parent_link, const std::string& lname,
                                                                         const UJoint* uj) -> int {
    Joint j;
    j.index = (int)joints.size();
    j.first_var = joints.empty() ? 0 : joints.back().first_var + joints.back().n_vars;
    Mat4 origin = identity();
    if (uj)
    {
      j.name = uj->name;
      const std::string& t = uj->type;
      j.type = t == "revolute" || t == "continuous" ? REVOLUTE :
               t == "prismatic" ? PRISMATIC : t == "floating" ? FLOATING : t == "planar" ? PLANAR : FIXED;
      if (j.type == REVOLUTE)
      {  // RevoluteJointModel::setAxis normalises (revolute_joint_model.cpp:73-82)
        const double n = sqrt(uj->axis[0] * uj->axis[0] + uj->axis[1] * uj->axis[1] + uj->axis[2] * uj->axis[2]);
        for (int k = 0; k < 3; ++k)
          j.axis[k] = uj->axis[k] / n;
      }
      else
        for (int k = 0; k < 3; ++k)
          j.axis[k] = uj->axis[k];  // PrismaticJointModel::setAxis keeps the raw axis (prismatic_joint_model.hpp:77-80)
      if (j.type == REVOLUTE || j.type == PRISMATIC)
      {  // jointBoundsFromURDF (robot_model.cpp:895-926)
        bool bounded = false;
        if (uj->has_safety)
        {
          bounded = true;
          j.lo = uj->soft_lower;
          j.hi = uj->soft_upper;
          if (uj->has_limit)
          {
            j.lo = std::max(j.lo, uj->lower);
            j.hi = std::min(j.hi, uj->upper);
          }
        }
        else if (uj->has_limit)
        {
          bounded = true;
          j.lo = uj->lower;
          j.hi = uj->upper;
        }
        if (t == "continuous" || (!bounded && j.type == REVOLUTE))
        {
          j.lo = -M_PI;
          j.hi = M_PI;
        }
        else if (!bounded)
        {
          j.lo = -std::numeric_limits<double>::max();
          j.hi = std::numeric_limits<double>::max();
        }
      }
      origin = poseToMatrix(uj->xyz, uj->rpy);
    }
    else
    {
      j.name = "ASSUMED_FIXED_ROOT_JOINT";
      j.type = FIXED;
      for (const auto& v : vjoints)
      {
        if (v.child != lname || v.frame.empty())
          continue;
        if (v.type == "fixed" || v.type == "planar" || v.type == "floating")
        {
          j.name = v.name;
          j.type = v.type == "fixed" ? FIXED : v.type == "planar" ? PLANAR : FLOATING;
          break;
        }
      }
    }
    j.n_vars = jointVarCount(j.type);
    j.parent_link = parent_link;
    Link l;
    l.index = (int)links.size();
    l.name = lname;
    l.parent = parent_link;
    l.shapes = ulinks[lname];
    l.origin = origin;
    j.child_link = l.index;
    const int ji = j.index, li = l.index;
    joint_index[j.name] = ji;
    link_index[lname] = li;
    joints.push_back(j);
    links.push_back(l);
    auto it = children.find(lname);
    if (it != children.end())
      for (const UJoint* cj : it->second)
      {
        const int c = build(li, cj->child, cj);
        links[li].child_joints.push_back(c);
      }
    return ji;
  };
  build(-1, roots[0], nullptr);
  const int n_vars = joints.back().first_var + joints.back().n_vars;

  // ---- buildMimic ---------------------------------------------------------------------------------------------
  for (auto& j : joints)
  {
    auto it = ujoints.find(j.name);
    if (it == ujoints.end() || !it->second.has_mimic)
      continue;
    auto jt = joint_index.find(it->second.mimic_joint);
    if (jt != joint_index.end() && joints[jt->second].n_vars == j.n_vars)
    {
      j.mimic = jt->second;
      j.mimic_factor = it->second.mimic_mult;
      j.mimic_offset = it->second.mimic_off;
    }
  }
  for (bool change = true; change;)
  {
    change = false;
    for (auto& j : joints)
    {
      if (j.mimic >= 0 && joints[j.mimic].mimic >= 0)
      {
        const Joint& m = joints[j.mimic];
        j.mimic_offset = j.mimic_offset + j.mimic_factor * m.mimic_offset;
        j.mimic_factor = j.mimic_factor * m.mimic_factor;
        j.mimic = m.mimic;
        change = true;
      }
      if (j.mimic == j.index)
      {
        for (auto& k : joints)
          k.mimic = -1;
        change = false;
        break;
      }
    }
  }
  for (auto& j : joints)
    if (j.mimic >= 0)
      joints[j.mimic].mimic_requests.push_back(j.index);

  // ---- computeDescendants (links only) -------------------------------------------------------------------------
  {
    std::vector<std::set<int>> desc(joints.size());
    std::set<int> seen;
    std::vector<int> parents;
    std::function<void(int)> helper = [&](int ji) {
      if (seen.count(ji))
        return;
      seen.insert(ji);
      const int lm = joints[ji].child_link;
      for (int p : parents)
        desc[p].insert(lm);
      desc[ji].insert(lm);
      parents.push_back(ji);
      for (int c : links[lm].child_joints)
        helper(c);
      for (int m : joints[ji].mimic_requests)
        helper(m);
      parents.pop_back();
    };
    helper(0);
    for (size_t i = 0; i < joints.size(); ++i)
      joints[i].descendant_links.assign(desc[i].begin(), desc[i].end());
  }

  // ---- default positions (robot_model.cpp:1428-1433) -------------------------------------------------------------
  std::vector<double> base(n_vars, 0.0);
  for (const auto& j : joints)
  {
    if (j.mimic >= 0 || j.n_vars == 0)
      continue;
    if (j.type == REVOLUTE || j.type == PRISMATIC)
      base[j.first_var] = (j.lo <= 0.0 && j.hi >= 0.0) ? 0.0 : (j.lo + j.hi) / 2.0;
    else if (j.type == FLOATING)
      base[j.first_var + 6] = 1.0;
  }
  for (const auto& j : joints)
    if (j.mimic >= 0)
      base[j.first_var] = base[joints[j.mimic].first_var] * j.mimic_factor + j.mimic_offset;

  // ---- buildGroups / addJointModelGroup ----------------------------------------------------------------------------
  std::map<std::string, Group> groups;
  {
    std::vector<bool> processed(gcfgs.size(), false);
    for (bool added = true; added;)
    {
      added = false;
      for (size_t i = 0; i < gcfgs.size(); ++i)
      {
        if (processed[i])
          continue;
        bool ok = true;
        for (const auto& s : gcfgs[i].subgroups)
          if (!groups.count(s))
            ok = false;
        if (!ok)
          continue;
        added = true;
        processed[i] = true;
        const GroupCfg& gc = gcfgs[i];
        if (groups.count(gc.name))
          continue;
        std::set<int> jset;
        for (const auto& ch : gc.chains)
        {
          if (!link_index.count(ch.first) || !link_index.count(ch.second))
            continue;
          const int base_l = link_index[ch.first];
          int lm = link_index[ch.second];
          std::vector<int> cj;
          while (lm >= 0)
          {
            if (lm == base_l)
              break;
            cj.push_back(lm);  // joint index == link index of its child
            lm = links[lm].parent;
          }
          if (lm != base_l)
          {
            lm = base_l;
            size_t index = 0;
            std::vector<int> cj2;
            while (lm >= 0)
            {
              for (size_t k = 0; k < cj.size(); ++k)
                if (cj[k] == lm)
                {
                  index = k + 1;
                  break;
                }
              if (index > 0)
                break;
              cj2.push_back(lm);
              lm = links[lm].parent;
            }
            if (index > 0)
            {
              jset.insert(cj.begin(), cj.begin() + index);
              jset.insert(cj2.begin(), cj2.end());
            }
          }
          else
            jset.insert(cj.begin(), cj.end());
        }
        for (const auto& jn : gc.joints)
          if (joint_index.count(jn))
            jset.insert(joint_index[jn]);
        for (const auto& ln : gc.links)
          if (link_index.count(ln))
            jset.insert(link_index[ln]);  // parent joint of link l has index l
        for (const auto& sg : gc.subgroups)
          if (groups.count(sg))
            jset.insert(groups[sg].joints.begin(), groups[sg].joints.end());
        if (jset.empty())
          continue;
        Group g;
        g.name = gc.name;
        g.joints.assign(jset.begin(), jset.end());
        groups[gc.name] = g;
      }
    }
  }
  auto git = groups.find(group_name);
  if (git == groups.end())
  {
    err = "no JointModelGroup named '" + group_name + "'";
    return false;
  }
  const Group& grp = git->second;
  std::set<int> in_group(grp.joints.begin(), grp.joints.end());
  std::function<bool(int)> includesParent = [&](int ji) -> bool {
    while (joints[ji].parent_link >= 0)
    {
      ji = joints[ji].parent_link;  // parent link's parent joint has the same index
      const Joint& j = joints[ji];
      if (in_group.count(ji) && j.n_vars > 0 && j.mimic < 0)
        return true;
      if (j.mimic >= 0)
      {
        const Joint& m = joints[j.mimic];
        if ((in_group.count(j.mimic) && m.n_vars > 0 && m.mimic < 0) || includesParent(j.mimic))
          return true;
      }
    }
    return false;
  };
  robot = HostRobot();
  robot.n_links = (int)links.size();
  robot.n_vars = n_vars;
  for (const auto& l : links)
  {
    const Joint& j = joints[l.index];
    robot.parent.push_back(l.parent);
    robot.joint_type.push_back(j.type);
    robot.first_var.push_back(j.first_var);
    robot.joint_axis.insert(robot.joint_axis.end(), j.axis, j.axis + 3);
    robot.joint_origin.insert(robot.joint_origin.end(), l.origin.begin(), l.origin.end());
    robot.link_names.push_back(l.name);
  }
  robot.base_positions = base;
  std::vector<int> roots_j;
  for (int ji : grp.joints)
  {
    const Joint& j = joints[ji];
    if (j.n_vars == 0)
      continue;
    for (int k = 0; k < j.n_vars; ++k)
      robot.group_var_index.push_back(j.first_var + k);
    if (j.type == REVOLUTE || j.type == PRISMATIC)
    {
      robot.group_lower.push_back(j.lo);
      robot.group_upper.push_back(j.hi);
    }
    else if (j.type == PLANAR)
    {  // unbounded x/y in MoveIt; the benchmark samples a 1 m box and full yaw
      robot.group_lower.insert(robot.group_lower.end(), { -0.5, -0.5, -M_PI });
      robot.group_upper.insert(robot.group_upper.end(), { 0.5, 0.5, M_PI });
    }
    else
      for (int k = 0; k < j.n_vars; ++k)
      {
        robot.group_lower.push_back(k < 3 ? -0.5 : -1.0);
        robot.group_upper.push_back(k < 3 ? 0.5 : 1.0);
      }
    if (j.mimic >= 0)
    {
      robot.mimic_dst.push_back(j.first_var);
      robot.mimic_src.push_back(joints[j.mimic].first_var);
      robot.mimic_factor.push_back(j.mimic_factor);
      robot.mimic_offset.push_back(j.mimic_offset);
    }
    else if (!includesParent(ji))
      roots_j.push_back(ji);
  }
  std::set<int> updated;
  for (int r : roots_j)
    updated.insert(joints[r].descendant_links.begin(), joints[r].descendant_links.end());

  // ---- BodyDecomposition per link with geometry (collision_distance_field_types.cpp:292-335, padding 0) --------------------
  // Spheres, cylinders and boxes; the bodies (containsPoint, bounding sphere / cylinder) are restated from geometric_shapes'
  // published bodies.cpp, which the reference does not vendor: parity unpinned for cylinders and boxes.
  struct Decomp { std::vector<std::array<double, 4>> spheres; std::vector<double> points; double bs[4] = { 0, 0, 0, 0 }; };
  std::map<int, Decomp> decomp;
  for (const auto& l : links)
  {
    if (l.shapes.empty())
      continue;
    Decomp d;
    bool first = true;
    for (const auto& s : l.shapes)
    {
      if (s.kind != "sphere" && s.kind != "cylinder" && s.kind != "box")
      {
        err = "link '" + l.name + "' has <" + s.kind +
              "> collision geometry; the URDF path decomposes <sphere>, <cylinder> and <box> (pass explicit spheres through "
              "b200sv_create for other shapes, as CollisionEnvDistanceField's link_body_decompositions argument does)";
        return false;
      }
      const double cx = Mc(s.origin, 0, 3), cy = Mc(s.origin, 1, 3), cz = Mc(s.origin, 2, 3);
      const double res = resolution;
      double r;  // radius of the body's bounding sphere (computeBoundingSphere)
      double half[3] = { 0, 0, 0 }, radius2 = 0.0, length2 = 0.0;
      if (s.kind == "sphere")
      {
        r = s.radius;
        d.spheres.push_back({ cx, cy, cz, r });  // collision_distance_field_types.cpp:63-67
      }
      else
      {
        // determineCollisionSpheres (:68-82) on the bounding cylinder (axis = its local z)
        Mat4 cp = s.origin;
        double cr, cl;
        if (s.kind == "cylinder")
        {
          radius2 = s.dims[0] * s.dims[0];
          length2 = 1.0 * s.dims[1] / 2.0;
          r = sqrt(length2 * length2 + radius2);
          cr = s.dims[0];
          cl = 2.0 * length2;
        }
        else
        {
          for (int k = 0; k < 3; ++k)
            half[k] = s.dims[k] * (1.0 / 2.0);
          r = sqrt(half[0] * half[0] + half[1] * half[1] + half[2] * half[2]);
          const double ang = 90.0 * (M_PI / 180.0), c = cos(ang), sn = sin(ang);
          Mat4 rot = identity();
          double a, b;
          if (half[0] > half[1] && half[0] > half[2])
          {
            cl = half[0] * 2.0; a = half[1]; b = half[2];
            M(rot, 0, 0) = c; M(rot, 0, 2) = sn; M(rot, 2, 0) = -sn; M(rot, 2, 2) = c;  // AngleAxis(90 deg, UnitY)
            cp = mulAffine(s.origin, rot);
          }
          else if (half[1] > half[2])
          {
            cl = half[1] * 2.0; a = half[2]; b = half[0];
            M(rot, 1, 1) = c; M(rot, 1, 2) = -sn; M(rot, 2, 1) = sn; M(rot, 2, 2) = c;  // AngleAxis(90 deg, UnitX)
            cp = mulAffine(s.origin, rot);
          }
          else
          {
            cl = half[2] * 2.0; a = half[1]; b = half[0];
          }
          cr = sqrt(a * a + b * b);
        }
        const long num_points = (long)ceil(cl / (cr / 2.0));
        const double spacing = cl / ((num_points * 1.0) - 1.0);
        for (long i = 1; i < num_points - 1; ++i)
        {
          const double z = (-cl / 2.0) + i * spacing;
          d.spheres.push_back({ Mc(cp, 0, 2) * z + Mc(cp, 0, 3), Mc(cp, 1, 2) * z + Mc(cp, 1, 3), Mc(cp, 2, 2) * z + Mc(cp, 2, 3), cr });
        }
      }
      // findInternalPointsConvex (find_internal_points.cpp:39-64) around the bounding sphere, containsPoint per body type
      const double xs = std::floor((cx - r - res) / res) * res, ys = std::floor((cy - r - res) / res) * res,
                   zs = std::floor((cz - r - res) / res) * res;
      const double xe = cx + r + res, ye = cy + r + res, ze = cz + r + res;
      for (double px = xs; px <= xe; px += res)
        for (double py = ys; py <= ye; py += res)
          for (double pz = zs; pz <= ze; pz += res)
          {
            bool inside;
            if (s.kind == "sphere")
            {
              const double dx = cx - px, dy = cy - py, dz = cz - pz;
              inside = dx * dx + dy * dy + dz * dz <= r * r;
            }
            else
            {
              const double vx = px - cx, vy = py - cy, vz = pz - cz;
              double al[3];
              for (int k = 0; k < 3; ++k)  // v . column k of the body's rotation
                al[k] = vx * Mc(s.origin, 0, k) + vy * Mc(s.origin, 1, k) + vz * Mc(s.origin, 2, k);
              if (s.kind == "box")
                inside = fabs(al[0]) <= half[0] && fabs(al[1]) <= half[1] && fabs(al[2]) <= half[2];
              else
              {
                const double remaining = radius2 - al[0] * al[0];
                inside = !(fabs(al[2]) > length2) && !(remaining < 0.0) && al[1] * al[1] <= remaining;
              }
            }
            if (inside)
            {
              d.points.push_back(px);
              d.points.push_back(py);
              d.points.push_back(pz);
            }
          }
      // bodies::mergeBoundingSpheres (geometric_shapes, restated; parity unpinned)
      if (first)
      {
        d.bs[0] = cx; d.bs[1] = cy; d.bs[2] = cz; d.bs[3] = r;
        first = false;
      }
      else if (r > 0.0)
      {
        const double ex = cx - d.bs[0], ey = cy - d.bs[1], ez = cz - d.bs[2];
        const double dist = sqrt(ex * ex + ey * ey + ez * ez);
        if (dist + d.bs[3] <= r)
        {
          d.bs[0] = cx; d.bs[1] = cy; d.bs[2] = cz; d.bs[3] = r;
        }
        else if (dist + r > d.bs[3])
        {
          const double fx = d.bs[0] - cx, fy = d.bs[1] - cy, fz = d.bs[2] - cz;
          const double n = sqrt(fx * fx + fy * fy + fz * fz);
          const double nr = (n + r + d.bs[3]) / 2.0;
          d.bs[0] = fx / n * (nr - r) + cx;
          d.bs[1] = fy / n * (nr - r) + cy;
          d.bs[2] = fz / n * (nr - r) + cz;
          d.bs[3] = nr;
        }
      }
    }
    decomp[l.index] = d;
  }

  // ---- DistanceFieldCacheEntry masks (collision_env_distance_field.cpp:735-838) ----------------------------------------
  std::map<std::string, std::map<std::string, bool>> acm;  // collision_matrix.cpp:67-78
  for (const auto& p : enabled)
    acm[p.first][p.second] = acm[p.second][p.first] = false;
  for (const auto& p : disabled)
    acm[p.first][p.second] = acm[p.second][p.first] = true;
  auto always = [&](const std::string& a, const std::string& b) {
    auto i = acm.find(a);
    if (i == acm.end())
      return false;
    auto k = i->second.find(b);
    return k != i->second.end() && k->second;
  };
  cm = HostCollisionModel();
  std::vector<int> gl(updated.begin(), updated.end());
  const int n = (int)gl.size();
  cm.n_glinks = n;
  cm.glink_index.assign(gl.begin(), gl.end());
  cm.self_enabled.assign(n, 1);
  cm.intra_enabled.assign((size_t)n * n, 0);
  cm.sphere_start.push_back(0);
  for (int i = 0; i < n; ++i)
  {
    const Link& li = links[gl[i]];
    auto d = decomp.find(gl[i]);
    if (d == decomp.end())
    {
      cm.self_enabled[i] = 0;
      cm.bounding_sphere.insert(cm.bounding_sphere.end(), { 0.0, 0.0, 0.0, 0.0 });
      cm.sphere_start.push_back(cm.sphere_start.back());
      continue;
    }
    if (always(li.name, li.name))
      cm.self_enabled[i] = 0;
    for (int j2 = 0; j2 < n; ++j2)
      cm.intra_enabled[(size_t)i * n + j2] = 1;
    for (int j2 = i + 1; j2 < n; ++j2)
      if (links[gl[j2]].name == li.name || always(li.name, links[gl[j2]].name))
        cm.intra_enabled[(size_t)i * n + j2] = 0;
    for (const auto& s : d->second.spheres)
      cm.sphere_xyzr.insert(cm.sphere_xyzr.end(), s.begin(), s.end());
    cm.bounding_sphere.insert(cm.bounding_sphere.end(), d->second.bs, d->second.bs + 4);
    cm.sphere_start.push_back(cm.sphere_start.back() + (int)d->second.spheres.size());
  }

  // ---- self-field points: links with geometry outside the updated set, posed at the base state (:907-953) ----------------
  self_local.clear();
  link_points.clear();
  for (const auto& l : links)
  {
    auto d = decomp.find(l.index);
    if (d == decomp.end())
      continue;
    link_points.push_back({ l.index, d->second.points });
    if (!updated.count(l.index))
      self_local.push_back({ l.index, d->second.points });
  }
  poseSelfPoints(robot, base, self_local, self_points);
  return true;
}

void poseSelfPoints(const HostRobot& robot, const std::vector<double>& positions, const SelfLocalPoints& self_local,
                    std::vector<double>& self_points)
{
  std::vector<Mat4> T;
  hostForwardKinematics(robot, positions, T);
  self_points.clear();
  for (const auto& lp : self_local)
  {
    const Mat4& Tm = T[lp.first];
    const auto& P = lp.second;
    for (size_t k = 0; k + 2 < P.size(); k += 3)
      for (int r = 0; r < 3; ++r)
      {
        double s = Mc(Tm, r, 0) * P[k];
        s += Mc(Tm, r, 1) * P[k + 1];
        s += Mc(Tm, r, 2) * P[k + 2];
        self_points.push_back(s + Mc(Tm, r, 3));
      }
  }
}
}  // namespace b200sv
