// See collision_env_b200.hpp.  Everything numerical happens behind the C-ABI; this file only flattens MoveIt's
// objects into the plain arrays of include/b200sv.h, exactly where CollisionEnvDistanceField builds its own
// DistanceFieldCacheEntry / GroupStateRepresentation (collision_env_distance_field.cpp:710-958, 1157-1252).
#include "collision_env_b200.hpp"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include <moveit/robot_model/robot_model.hpp>
#include <moveit/robot_state/robot_state.hpp>
#include <moveit/utils/logger.hpp>
#include <geometric_shapes/shapes.h>

namespace collision_detection
{
namespace
{
rclcpp::Logger getLogger()
{
  return moveit::getLogger("moveit.core.collision_env_b200");
}
const double STATE_EPSILON = 0.001;  // EPSILON of collision_env_distance_field.cpp:55

b200sv_field_cfg fieldCfg(const Eigen::Vector3d& size, const Eigen::Vector3d& origin, double res, double max_prop,
                          double tol, bool use_signed_distance_field)
{
  b200sv_field_cfg f{};
  for (int k = 0; k < 3; ++k)
  {
    f.size[k] = size[k];
    f.origin[k] = origin[k];
  }
  f.resolution = res;
  f.max_propagation_distance = max_prop;
  f.collision_tolerance = tol;
  f.use_signed_distance_field = use_signed_distance_field ? 1 : 0;
  return f;
}
}  // namespace

CollisionEnvB200::CollisionEnvB200(const moveit::core::RobotModelConstPtr& robot_model, double padding, double scale)
  : CollisionEnv(robot_model, padding, scale)
{
  initialize();
  observer_handle_ = getWorld()->addObserver(
      [this](const World::ObjectConstPtr& object, World::Action action) { notifyObjectChange(object, action); });
}

CollisionEnvB200::CollisionEnvB200(const moveit::core::RobotModelConstPtr& robot_model, const WorldPtr& world,
                                   double padding, double scale)
  : CollisionEnv(robot_model, world, padding, scale)
{
  initialize();
  observer_handle_ = getWorld()->addObserver(
      [this](const World::ObjectConstPtr& object, World::Action action) { notifyObjectChange(object, action); });
  getWorld()->notifyObserverAllObjects(observer_handle_, World::CREATE);
}

CollisionEnvB200::CollisionEnvB200(const CollisionEnvB200& other, const WorldPtr& world) : CollisionEnv(other, world)
{
  size_ = other.size_;
  origin_ = other.origin_;
  resolution_ = other.resolution_;
  collision_tolerance_ = other.collision_tolerance_;
  max_propogation_distance_ = other.max_propogation_distance_;
  reference_propagation_ = other.reference_propagation_;
  link_body_decomposition_vector_ = other.link_body_decomposition_vector_;
  link_body_decomposition_index_map_ = other.link_body_decomposition_index_map_;
  // contexts are per (group, non-group state, ACM) and own device memory: a child scene builds its own lazily and
  // replays its world into them, as the reference rebuilds its world field (:111, :118)
  observer_handle_ = getWorld()->addObserver(
      [this](const World::ObjectConstPtr& object, World::Action action) { notifyObjectChange(object, action); });
  getWorld()->notifyObserverAllObjects(observer_handle_, World::CREATE);
}

CollisionEnvB200::GroupContext::~GroupContext()
{
  if (ctx)
    b200sv_destroy(ctx);  // runs when the last holder drops the shared_ptr: never under a thread that is still checking
}

CollisionEnvB200::~CollisionEnvB200()
{
  getWorld()->removeObserver(observer_handle_);
}

void CollisionEnvB200::initialize()
{
  size_ = Eigen::Vector3d(DEFAULT_SIZE_X, DEFAULT_SIZE_Y, DEFAULT_SIZE_Z);
  origin_ = Eigen::Vector3d(0, 0, 0);
  resolution_ = DEFAULT_RESOLUTION;
  collision_tolerance_ = DEFAULT_COLLISION_TOLERANCE;
  max_propogation_distance_ = DEFAULT_MAX_PROPOGATION_DISTANCE;
  if (const char* mode = std::getenv("B200SV_PROPAGATION"))
    reference_propagation_ = std::string(mode) == "reference";
  // addLinkBodyDecompositions (collision_env_distance_field.cpp:960-979): the reference's own BodyDecomposition,
  // so spheres / interior points / bounding spheres of meshes, boxes and cylinders are the reference's, bit for bit
  for (const moveit::core::LinkModel* link : robot_model_->getLinkModelsWithCollisionGeometry())
  {
    if (link->getShapes().empty())
      continue;
    link_body_decomposition_vector_.push_back(std::make_shared<const BodyDecomposition>(
        link->getShapes(), link->getCollisionOriginTransforms(), resolution_, getLinkPadding(link->getName())));
    link_body_decomposition_index_map_[link->getName()] = link_body_decomposition_vector_.size() - 1;
  }
}

// ---- generateDistanceFieldCacheEntry, the part that depends on (state, ACM): enable masks + attached bodies (:735-866) ----
void CollisionEnvB200::flattenModel(const moveit::core::JointModelGroup* jmg, const moveit::core::RobotState& state,
                                    const AllowedCollisionMatrix* acm, FlatModel& out) const
{
  const std::vector<const moveit::core::LinkModel*>& glinks = jmg->getUpdatedLinkModels();
  const int n = static_cast<int>(glinks.size());
  out.glink_index.assign(n, 0);
  out.sphere_start.assign(1, 0);
  out.self_enabled.assign(n, 1);
  out.intra_enabled.assign(static_cast<std::size_t>(n) * n, 0);
  out.sphere_xyzr.clear();
  out.bounding_sphere.assign(4 * static_cast<std::size_t>(n), 0.0);
  std::vector<uint8_t>& self_enabled = out.self_enabled;
  std::vector<uint8_t>& intra_enabled = out.intra_enabled;
  for (int i = 0; i < n; ++i)
  {
    const std::string& name = glinks[i]->getName();
    out.glink_index[i] = static_cast<int32_t>(glinks[i]->getLinkIndex());
    auto bd = link_body_decomposition_index_map_.find(name);
    if (glinks[i]->getShapes().empty() || bd == link_body_decomposition_index_map_.end())
    {
      self_enabled[i] = 0;
      out.sphere_start.push_back(out.sphere_start.back());
      continue;
    }
    AllowedCollision::Type t;
    if (acm && acm->getEntry(name, name, t) && t == AllowedCollision::ALWAYS)
      self_enabled[i] = 0;
    for (int j = 0; j < n; ++j)
      intra_enabled[static_cast<std::size_t>(i) * n + j] = 1;
    for (int j = i + 1; j < n; ++j)
      if (name == glinks[j]->getName() ||
          (acm && acm->getEntry(name, glinks[j]->getName(), t) && t == AllowedCollision::ALWAYS))
        intra_enabled[static_cast<std::size_t>(i) * n + j] = 0;
    const BodyDecompositionConstPtr& d = link_body_decomposition_vector_[bd->second];
    for (const CollisionSphere& cs : d->getCollisionSpheres())
      out.sphere_xyzr.insert(out.sphere_xyzr.end(),
                             { cs.relative_vec_.x(), cs.relative_vec_.y(), cs.relative_vec_.z(), cs.radius_ });
    out.sphere_start.push_back(static_cast<int32_t>(out.sphere_xyzr.size() / 4));
    const bodies::BoundingSphere& bs = d->getRelativeBoundingSphere();
    out.bounding_sphere[4 * i + 0] = bs.center.x();
    out.bounding_sphere[4 * i + 1] = bs.center.y();
    out.bounding_sphere[4 * i + 2] = bs.center.z();
    out.bounding_sphere[4 * i + 3] = bs.radius;
  }
  // ---- attached bodies: dfce->attached_body_names_ follow the links (:798-822, 840-866) -------------------------------
  // One entry per SHAPE of each body attached to a group link; the entry repeats the link's index (that is how
  // libb200sv recognises it), its spheres and bounding sphere are expressed in the LINK frame
  // (AttachedBody::getShapePosesInLinkFrame() * BodyDecomposition-relative coordinates).  Masks as the reference sets
  // them: (link, body) is disabled by an ALWAYS entry or a touch link ONLY for the link the body hangs on; body-body by
  // an ALWAYS entry; shapes of one body never test each other.
  struct BodyEntry
  {
    int link_i;  // index into glinks
    int body;    // which attached body (its shapes share the id: b200sv_collision_model::body_id)
    std::string name;
    std::vector<double> xyzr;
    double bs[4];
  };
  std::vector<BodyEntry> body_entries;
  std::string& sig = out.attached_signature;
  sig.clear();
  char num[64];
  int n_bodies = 0;
  for (int i = 0; i < n; ++i)
  {
    std::vector<const moveit::core::AttachedBody*> attached;
    state.getAttachedBodies(attached, glinks[i]);
    for (const moveit::core::AttachedBody* att : attached)
    {
      // what compareCacheEntryToState looks at for attached bodies (:1293-1312): which bodies, on which links, their
      // touch links and shapes -- plus the shape poses in the link frame, which the baked-in spheres depend on
      sig += glinks[i]->getName() + "<" + att->getName() + ":";
      for (const std::string& tl : att->getTouchLinks())
        sig += tl + ",";
      const EigenSTL::vector_Isometry3d& poses = att->getShapePosesInLinkFrame();
      for (std::size_t k = 0; k < att->getShapes().size(); ++k)
      {
        const BodyDecomposition d(att->getShapes()[k], resolution_, 0.0);
        BodyEntry e;
        e.link_i = i;
        e.body = n_bodies;
        e.name = att->getName();
        for (const CollisionSphere& cs : d.getCollisionSpheres())
        {
          const Eigen::Vector3d c = poses[k] * cs.relative_vec_;
          e.xyzr.insert(e.xyzr.end(), { c.x(), c.y(), c.z(), cs.radius_ });
        }
        const Eigen::Vector3d bc = poses[k] * d.getRelativeBoundingSphere().center;
        e.bs[0] = bc.x(); e.bs[1] = bc.y(); e.bs[2] = bc.z(); e.bs[3] = d.getRelativeBoundingSphere().radius;
        snprintf(num, sizeof(num), "|%d;%zu;", static_cast<int>(att->getShapes()[k]->type), e.xyzr.size());
        sig += num;
        for (int r = 0; r < 3; ++r)
          for (int col = 0; col < 4; ++col)
          {
            snprintf(num, sizeof(num), "%.9g,", poses[k](r, col));
            sig += num;
          }
        body_entries.push_back(std::move(e));
      }
      sig += ">";
      ++n_bodies;
    }
  }
  out.entry_names.clear();
  out.entry_is_attached.clear();
  out.body_id.assign(n, -1);
  for (int i = 0; i < n; ++i)
  {
    out.entry_names.push_back(glinks[i]->getName());
    out.entry_is_attached.push_back(false);
  }
  if (!body_entries.empty())
  {
    const int nb = static_cast<int>(body_entries.size()), nt = n + nb;
    std::vector<uint8_t> intra2(static_cast<std::size_t>(nt) * nt, 0);
    for (int i = 0; i < n; ++i)
    {
      for (int j = 0; j < n; ++j)
        intra2[static_cast<std::size_t>(i) * nt + j] = intra_enabled[static_cast<std::size_t>(i) * n + j];
      if (!glinks[i]->getShapes().empty())
        for (int b = 0; b < nb; ++b)
          intra2[static_cast<std::size_t>(i) * nt + n + b] = 1;  // all_true
    }
    for (int b = 0; b < nb; ++b)
    {
      const BodyEntry& e = body_entries[b];
      const moveit::core::AttachedBody* att = state.getAttachedBody(e.name);
      const std::string& link_name = glinks[e.link_i]->getName();
      AllowedCollision::Type t;
      if ((acm && acm->getEntry(link_name, e.name, t) && t == AllowedCollision::ALWAYS) ||
          att->getTouchLinks().count(link_name))
        intra2[static_cast<std::size_t>(e.link_i) * nt + n + b] = 0;
      out.glink_index.push_back(static_cast<int32_t>(glinks[e.link_i]->getLinkIndex()));
      self_enabled.push_back(!(acm && acm->getEntry(e.name, e.name, t) && t == AllowedCollision::ALWAYS));
      for (int b2 = b + 1; b2 < nb; ++b2)
        intra2[static_cast<std::size_t>(n + b) * nt + n + b2] =
            !(body_entries[b2].name == e.name ||
              (acm && acm->getEntry(e.name, body_entries[b2].name, t) && t == AllowedCollision::ALWAYS));
      out.sphere_xyzr.insert(out.sphere_xyzr.end(), e.xyzr.begin(), e.xyzr.end());
      out.sphere_start.push_back(static_cast<int32_t>(out.sphere_xyzr.size() / 4));
      out.bounding_sphere.insert(out.bounding_sphere.end(), e.bs, e.bs + 4);
      out.entry_names.push_back(e.name);
      out.entry_is_attached.push_back(true);
      out.body_id.push_back(e.body);
    }
    intra_enabled.swap(intra2);
  }
}

// ---- the context itself: RobotModel + group layout + collision model -> b200sv_create; self-field link points; world ----
std::shared_ptr<CollisionEnvB200::GroupContext>
CollisionEnvB200::createContext(const moveit::core::JointModelGroup* jmg, const moveit::core::RobotState& state,
                                const FlatModel& fm) const
{
  // b200sv_robot: RobotModel in link-index order + the group's variable layout
  const auto& links = robot_model_->getLinkModels();
  const int n_links = static_cast<int>(links.size());
  std::vector<int32_t> parent(n_links), joint_type(n_links), first_var(n_links);
  std::vector<double> axis(3 * n_links, 0.0), origin(16 * n_links);
  for (const moveit::core::LinkModel* l : links)
  {
    const int i = static_cast<int>(l->getLinkIndex());
    const moveit::core::JointModel* j = l->getParentJointModel();
    parent[i] = l->getParentLinkModel() ? static_cast<int32_t>(l->getParentLinkModel()->getLinkIndex()) : -1;
    first_var[i] = j->getVariableCount() ? static_cast<int32_t>(j->getFirstVariableIndex()) : 0;
    Eigen::Vector3d a(1, 0, 0);
    switch (j->getType())
    {
      case moveit::core::JointModel::REVOLUTE:
        joint_type[i] = B200SV_JOINT_REVOLUTE;
        a = static_cast<const moveit::core::RevoluteJointModel*>(j)->getAxis();
        break;
      case moveit::core::JointModel::PRISMATIC:
        joint_type[i] = B200SV_JOINT_PRISMATIC;
        a = static_cast<const moveit::core::PrismaticJointModel*>(j)->getAxis();
        break;
      case moveit::core::JointModel::PLANAR: joint_type[i] = B200SV_JOINT_PLANAR; break;
      case moveit::core::JointModel::FLOATING: joint_type[i] = B200SV_JOINT_FLOATING; break;
      default: joint_type[i] = B200SV_JOINT_FIXED; break;
    }
    for (int k = 0; k < 3; ++k)
      axis[3 * i + k] = a[k];
    Eigen::Map<Eigen::Matrix4d>(origin.data() + 16 * i) = l->getJointOriginTransform().matrix();  // column-major
  }
  const std::vector<int>& gvi = jmg->getVariableIndexList();
  std::vector<int32_t> group_var_index(gvi.begin(), gvi.end());
  std::vector<int32_t> mimic_dst, mimic_src;
  std::vector<double> mimic_factor, mimic_offset;
  for (const moveit::core::JointModel* jm : jmg->getMimicJointModels())  // RobotState::updateMimicJoints(group)
  {
    mimic_dst.push_back(static_cast<int32_t>(jm->getFirstVariableIndex()));
    mimic_src.push_back(static_cast<int32_t>(jm->getMimic()->getFirstVariableIndex()));
    mimic_factor.push_back(jm->getMimicFactor());
    mimic_offset.push_back(jm->getMimicOffset());
  }
  b200sv_robot r{};
  r.n_links = n_links;
  r.n_vars = static_cast<int32_t>(robot_model_->getVariableCount());
  r.parent = parent.data();
  r.joint_type = joint_type.data();
  r.joint_axis = axis.data();
  r.joint_origin = origin.data();
  r.first_var = first_var.data();
  r.base_positions = state.getVariablePositions();
  r.n_group_vars = static_cast<int32_t>(group_var_index.size());
  r.group_var_index = group_var_index.data();
  r.n_mimic = static_cast<int32_t>(mimic_dst.size());
  r.mimic_dst = mimic_dst.data();
  r.mimic_src = mimic_src.data();
  r.mimic_factor = mimic_factor.data();
  r.mimic_offset = mimic_offset.data();

  b200sv_collision_model m{};
  m.n_glinks = static_cast<int32_t>(fm.glink_index.size());
  m.glink_index = fm.glink_index.data();
  m.self_enabled = fm.self_enabled.data();
  m.intra_enabled = fm.intra_enabled.data();
  m.sphere_start = fm.sphere_start.data();
  m.sphere_xyzr = fm.sphere_xyzr.data();
  m.bounding_sphere = fm.bounding_sphere.data();
  m.body_id = fm.body_id.data();  // multi-shape bodies are culled as a whole, as the reference does (:455-507)

  const b200sv_field_cfg f = fieldCfg(size_, origin_, resolution_, max_propogation_distance_, collision_tolerance_,
                                      use_signed_distance_field_);
  auto gc = std::make_shared<GroupContext>();
  int device = 0;  // one process per GPU: honour CUDA_VISIBLE_DEVICES / the launcher's LOCAL_RANK
  if (const char* lr = std::getenv("LOCAL_RANK"))
    device = std::atoi(lr);
  if (b200sv_create(device, &r, &m, &f, &gc->ctx) != B200SV_OK)
  {
    RCLCPP_ERROR(getLogger(), "b200sv_create failed: %s", b200sv_last_error(nullptr));
    return nullptr;
  }
  if (reference_propagation_ &&
      (b200sv_field_set_propagation(gc->ctx, B200SV_FIELD_WORLD, B200SV_PROPAGATION_REFERENCE) != B200SV_OK ||
       b200sv_field_set_propagation(gc->ctx, B200SV_FIELD_SELF, B200SV_PROPAGATION_REFERENCE) != B200SV_OK))
  {
    RCLCPP_ERROR(getLogger(), "b200sv_field_set_propagation failed: %s", b200sv_last_error(gc->ctx));
    return nullptr;
  }
  // non-group variables to watch (dfce->state_check_indices_, :886-896)
  std::vector<bool> in_group(robot_model_->getVariableCount(), false);
  for (const moveit::core::JointModel* jm : jmg->getActiveJointModels())
    for (std::size_t k = 0; k < jm->getVariableCount(); ++k)
      in_group[jm->getFirstVariableIndex() + k] = true;
  for (std::size_t v = 0; v < in_group.size(); ++v)
    if (!in_group[v])
    {
      gc->non_group_indices.push_back(static_cast<int>(v));
      gc->non_group_values.push_back(state.getVariablePosition(static_cast<int>(v)));
    }
  // self field (:907-953): the interior points of the links the group does not update go to the library ONCE, in the link
  // frame; b200sv_set_base_state poses them at the cached state and rebuilds the self field -- now and whenever a joint
  // outside the group moves
  for (const moveit::core::LinkModel* lm : robot_model_->getLinkModelsWithCollisionGeometry())
  {
    if (jmg->isLinkUpdated(lm->getName()))
      continue;
    auto bd = link_body_decomposition_index_map_.find(lm->getName());
    if (bd == link_body_decomposition_index_map_.end())
      continue;
    const EigenSTL::vector_Vector3d& pts = link_body_decomposition_vector_[bd->second]->getCollisionPoints();
    std::vector<double> flat;
    flat.reserve(3 * pts.size());
    for (const Eigen::Vector3d& p : pts)
      flat.insert(flat.end(), { p.x(), p.y(), p.z() });
    if (b200sv_set_self_link_points(gc->ctx, static_cast<int32_t>(lm->getLinkIndex()), flat.data(), pts.size()) != B200SV_OK)
    {
      RCLCPP_ERROR(getLogger(), "b200sv_set_self_link_points failed: %s", b200sv_last_error(gc->ctx));
      return nullptr;
    }
  }
  if (b200sv_set_base_state(gc->ctx, state.getVariablePositions(), 1) != B200SV_OK)
  {
    RCLCPP_ERROR(getLogger(), "b200sv_set_base_state failed: %s", b200sv_last_error(gc->ctx));
    return nullptr;
  }
  gc->self_enabled = fm.self_enabled;
  gc->intra_enabled = fm.intra_enabled;
  gc->attached_signature = fm.attached_signature;
  gc->entry_names = fm.entry_names;
  gc->entry_is_attached = fm.entry_is_attached;
  gc->sphere_start = fm.sphere_start;
  gc->body_id = fm.body_id;
  gc->sphere_radii.resize(fm.sphere_xyzr.size() / 4);
  for (std::size_t k = 0; k < gc->sphere_radii.size(); ++k)
    gc->sphere_radii[k] = fm.sphere_xyzr[4 * k + 3];
  for (int32_t li : fm.glink_index)
    gc->entry_joint_names.push_back(links[static_cast<std::size_t>(li)]->getParentJointModel()->getName());
  return gc;
}

// ---- getDistanceFieldCacheEntry (:177-233): reuse the cached entry when it matches (state outside the group, ACM, attached
// bodies), otherwise bring it up to date.  The reference regenerates the whole entry; here only what changed is re-derived:
//   attached bodies differ  -> a new context (the sphere tables change shape)
//   a non-group joint moved -> b200sv_set_base_state: constant transforms + self field
//   the ACM's masks differ  -> b200sv_set_acm_masks
std::shared_ptr<CollisionEnvB200::GroupContext>
CollisionEnvB200::getContext(const std::string& group_name, const moveit::core::RobotState& state,
                             const AllowedCollisionMatrix* acm, std::unique_lock<std::mutex>& lock) const
{
  const moveit::core::JointModelGroup* jmg = robot_model_->getJointModelGroup(group_name);
  if (!jmg)
  {
    RCLCPP_WARN(getLogger(), "No group %s", group_name.c_str());  // :716-720
    return nullptr;
  }
  FlatModel fm;
  flattenModel(jmg, state, acm, fm);
  std::shared_ptr<GroupContext> gc;
  std::size_t world_version = 0;
  {
    std::scoped_lock map_lock(mutex_);
    auto it = contexts_.find(group_name);
    if (it != contexts_.end())
      gc = it->second;
    world_version = world_version_;
  }
  if (gc)
  {
    lock = std::unique_lock<std::mutex>(gc->mutex);
    if (gc->attached_signature != fm.attached_signature)
    {  // compareCacheEntryToState, attached-body part (:1293-1312): different bodies -> a new entry
      lock.unlock();
      gc.reset();
    }
  }
  if (!gc)
  {
    gc = createContext(jmg, state, fm);
    if (!gc)
      return nullptr;
    lock = std::unique_lock<std::mutex>(gc->mutex);
    gc->world_version = static_cast<std::size_t>(-1);
    std::scoped_lock map_lock(mutex_);
    contexts_[group_name] = gc;  // the entry it replaces lives on until its last user is done
  }
  else
  {
    // compareCacheEntryToState (:1254-1291): any variable outside the group moved by more than EPSILON
    bool same = true;
    for (std::size_t i = 0; i < gc->non_group_indices.size() && same; ++i)
      same = std::fabs(state.getVariablePosition(gc->non_group_indices[i]) - gc->non_group_values[i]) <= STATE_EPSILON;
    if (!same)
    {
      if (b200sv_set_base_state(gc->ctx, state.getVariablePositions(), 1) != B200SV_OK)
      {
        RCLCPP_ERROR(getLogger(), "b200sv_set_base_state failed: %s", b200sv_last_error(gc->ctx));
        return nullptr;
      }
      for (std::size_t i = 0; i < gc->non_group_indices.size(); ++i)
        gc->non_group_values[i] = state.getVariablePosition(gc->non_group_indices[i]);
    }
    // compareCacheEntryToAllowedCollisionMatrix (:1314-1370): the entry's enable masks against the ones this ACM gives
    if (gc->self_enabled != fm.self_enabled || gc->intra_enabled != fm.intra_enabled)
    {
      if (b200sv_set_acm_masks(gc->ctx, fm.self_enabled.data(), fm.intra_enabled.data()) != B200SV_OK)
      {
        RCLCPP_ERROR(getLogger(), "b200sv_set_acm_masks failed: %s", b200sv_last_error(gc->ctx));
        return nullptr;
      }
      gc->self_enabled = fm.self_enabled;
      gc->intra_enabled = fm.intra_enabled;
    }
  }
  if (gc->world_version != world_version)
  {  // a context created after the world was filled, or one that missed updates: replay what the world holds
    if (!replayWorld(gc->ctx))
      return nullptr;
    gc->world_version = world_version;
  }
  return gc;
}

namespace
{
// Spheres, cylinders and boxes are decomposed on the device (same lattice and containsPoint as BodyDecomposition builds on the
// host with padding 0); everything else keeps the reference's host path.
bool devicePrimitive(const shapes::Shape* shape, const Eigen::Isometry3d& pose, b200sv_shape& out)
{
  out = b200sv_shape{};
  out.lattice_frame = B200SV_LATTICE_BODY;
  switch (shape->type)
  {
    case shapes::SPHERE:
      out.type = B200SV_SHAPE_SPHERE;
      out.dims[0] = static_cast<const shapes::Sphere*>(shape)->radius;
      break;
    case shapes::CYLINDER:
      out.type = B200SV_SHAPE_CYLINDER;
      out.dims[0] = static_cast<const shapes::Cylinder*>(shape)->radius;
      out.dims[1] = static_cast<const shapes::Cylinder*>(shape)->length;
      break;
    case shapes::BOX:
      out.type = B200SV_SHAPE_BOX;
      for (int k = 0; k < 3; ++k)
        out.dims[k] = static_cast<const shapes::Box*>(shape)->size[k];
      break;
    default:
      return false;
  }
  for (int r = 0; r < 3; ++r)
    for (int k = 0; k < 4; ++k)
      out.pose[4 * r + k] = pose(r, k);
  return true;
}
}  // namespace

void CollisionEnvB200::setReferencePropagation(bool on)
{
  {
    std::scoped_lock lock(mutex_);
    if (on == reference_propagation_)
      return;
    reference_propagation_ = on;
    contexts_.clear();  // their holders keep them alive until their checks return; successors are built in the new mode
  }
  // the world's records are re-derived in the new mode's form (device primitives, or the reference's points for every shape)
  getWorld()->notifyObserverAllObjects(observer_handle_, World::CREATE);
}

bool CollisionEnvB200::replayWorld(b200sv_ctx* ctx) const
{
  if (reference_propagation_)
  {  // object by object, in the order World::notifyObserverAllObjects walks its map: every addPointsToField call
     // propagates on top of the state the one before left, as in the reference (:1704-1779)
    std::vector<EigenSTL::vector_Vector3d> per_object;
    {
      std::scoped_lock map_lock(mutex_);
      for (const auto& kv : world_points_)
        per_object.push_back(kv.second.points);
    }
    bool ok = b200sv_field_reset(ctx, B200SV_FIELD_WORLD) == B200SV_OK;
    for (std::size_t i = 0; ok && i < per_object.size(); ++i)
      if (!per_object[i].empty())
        ok = b200sv_field_add_points(ctx, B200SV_FIELD_WORLD, per_object[i][0].data(), per_object[i].size()) == B200SV_OK;
    if (!ok)
      RCLCPP_ERROR(getLogger(), "replaying the world into the field failed: %s", b200sv_last_error(ctx));
    return ok;
  }
  EigenSTL::vector_Vector3d all;
  std::vector<b200sv_shape> shapes;
  {
    std::scoped_lock map_lock(mutex_);
    for (const auto& kv : world_points_)
    {
      all.insert(all.end(), kv.second.points.begin(), kv.second.points.end());
      shapes.insert(shapes.end(), kv.second.shapes.begin(), kv.second.shapes.end());
    }
  }
  // everything the world holds, ONE transform at the end
  bool ok = b200sv_field_reset(ctx, B200SV_FIELD_WORLD) == B200SV_OK &&
            b200sv_field_begin_update(ctx, B200SV_FIELD_WORLD) == B200SV_OK;
  if (ok && !all.empty())
    ok = b200sv_field_add_points(ctx, B200SV_FIELD_WORLD, all[0].data(), all.size()) == B200SV_OK;
  for (std::size_t i = 0; ok && i < shapes.size(); ++i)
    ok = b200sv_field_add_shape(ctx, B200SV_FIELD_WORLD, &shapes[i], 0) == B200SV_OK;
  ok = b200sv_field_end_update(ctx, B200SV_FIELD_WORLD) == B200SV_OK && ok;
  if (!ok)
    RCLCPP_ERROR(getLogger(), "replaying the world into the field failed: %s", b200sv_last_error(ctx));
  return ok;
}

void CollisionEnvB200::notifyObjectChange(const ObjectConstPtr& obj, World::Action action)
{
  // updateDistanceObject (:1730-1779): subtract what the object contributed, add its new decomposition
  WorldObjectRecord old_rec, new_rec;
  std::vector<std::shared_ptr<GroupContext>> live;
  if (action != World::DESTROY && getWorld()->hasObject(obj->id_))
  {
    World::ObjectConstPtr object = getWorld()->getObject(obj->id_);
    for (std::size_t i = 0; i < object->shapes_.size(); ++i)
    {
      const shapes::ShapeConstPtr& shape = object->shapes_[i];
      b200sv_shape prim;
      if (shape->type == shapes::OCTREE)
      {
        PosedBodyPointDecomposition pts(static_cast<const shapes::OcTree*>(shape.get())->octree);  // types.cpp:355-365
        new_rec.points.insert(new_rec.points.end(), pts.getCollisionPoints().begin(), pts.getCollisionPoints().end());
      }
      else if (!reference_propagation_ &&
               devicePrimitive(shape.get(), getWorld()->getGlobalShapeTransform(obj->id_, static_cast<int>(i)), prim))
        new_rec.shapes.push_back(prim);  // (reference mode: every shape takes the reference's own host decomposition below,
                                         //  so the points and their order are the reference's)
      else
      {
        BodyDecompositionConstPtr bd = std::make_shared<const BodyDecomposition>(shape, resolution_, 0.0);
        PosedBodyPointDecomposition pts(bd, getWorld()->getGlobalShapeTransform(obj->id_, static_cast<int>(i)));
        new_rec.points.insert(new_rec.points.end(), pts.getCollisionPoints().begin(), pts.getCollisionPoints().end());
      }
    }
  }
  std::size_t version_before = 0, version_after = 0;
  {
    std::scoped_lock lock(mutex_);
    auto cur = world_points_.find(obj->id_);
    if (cur != world_points_.end())
      std::swap(old_rec, cur->second);
    if (action != World::DESTROY && getWorld()->hasObject(obj->id_))
      world_points_[obj->id_] = new_rec;
    else
      world_points_.erase(obj->id_);
    version_before = world_version_;
    version_after = ++world_version_;
    for (auto& kv : contexts_)
      live.push_back(kv.second);
  }
  for (const std::shared_ptr<GroupContext>& gc : live)
  {
    std::scoped_lock ctx_lock(gc->mutex);
    if (gc->world_version != version_before)
      continue;  // it missed an earlier change: the next check replays the whole world into it
    b200sv_ctx* ctx = gc->ctx;
    if (reference_propagation_)
    {  // removePointsFromField, then addPointsToField (:1713-1725), each with its own propagation, as the reference
      bool ok = old_rec.shapes.empty();  // a record made before the mode was switched: replay instead
      if (ok && !old_rec.points.empty())
        ok = b200sv_field_remove_points(ctx, B200SV_FIELD_WORLD, old_rec.points[0].data(), old_rec.points.size()) == B200SV_OK;
      if (ok && !new_rec.points.empty())
        ok = b200sv_field_add_points(ctx, B200SV_FIELD_WORLD, new_rec.points[0].data(), new_rec.points.size()) == B200SV_OK;
      if (ok)
        gc->world_version = version_after;
      else
        RCLCPP_ERROR(getLogger(), "world update of the field failed (the next check replays the world): %s",
                     b200sv_last_error(ctx));
      continue;
    }
    // removePointsFromField, then addPointsToField (:1713-1725) -- as ONE rebuild of the field
    bool ok = b200sv_field_begin_update(ctx, B200SV_FIELD_WORLD) == B200SV_OK;
    if (ok && !old_rec.points.empty())
      ok = b200sv_field_remove_points(ctx, B200SV_FIELD_WORLD, old_rec.points[0].data(), old_rec.points.size()) == B200SV_OK;
    for (std::size_t i = 0; ok && i < old_rec.shapes.size(); ++i)
      ok = b200sv_field_add_shape(ctx, B200SV_FIELD_WORLD, &old_rec.shapes[i], 1) == B200SV_OK;
    if (ok && !new_rec.points.empty())
      ok = b200sv_field_add_points(ctx, B200SV_FIELD_WORLD, new_rec.points[0].data(), new_rec.points.size()) == B200SV_OK;
    for (std::size_t i = 0; ok && i < new_rec.shapes.size(); ++i)
      ok = b200sv_field_add_shape(ctx, B200SV_FIELD_WORLD, &new_rec.shapes[i], 0) == B200SV_OK;
    ok = b200sv_field_end_update(ctx, B200SV_FIELD_WORLD) == B200SV_OK && ok;
    if (ok)
      gc->world_version = version_after;
    else
      RCLCPP_ERROR(getLogger(), "world update of the field failed (the next check replays the world): %s",
                   b200sv_last_error(ctx));
  }
}

void CollisionEnvB200::setWorld(const WorldPtr& world)
{
  if (world == getWorld())
    return;
  getWorld()->removeObserver(observer_handle_);
  {
    std::scoped_lock lock(mutex_);
    world_points_.clear();
    ++world_version_;  // every context replays the (new) world at its next check
  }
  CollisionEnv::setWorld(world);
  observer_handle_ = getWorld()->addObserver(
      [this](const World::ObjectConstPtr& object, World::Action action) { notifyObjectChange(object, action); });
  getWorld()->notifyObserverAllObjects(observer_handle_, World::CREATE);
}

bool CollisionEnvB200::checkCollisionBatch(const CollisionRequest& req, const moveit::core::RobotState& base_state,
                                           const AllowedCollisionMatrix* acm, const double* q, std::size_t n_states,
                                           uint8_t* valid_bitmap, uint32_t flags) const
{
  // An engine failure must never validate a state: the bitmap is cleared first (every state "in collision") and only a
  // successful check overwrites it.
  std::fill(valid_bitmap, valid_bitmap + (n_states + 7) / 8, static_cast<uint8_t>(0));
  std::unique_lock<std::mutex> ctx_lock;
  std::shared_ptr<GroupContext> gc = getContext(req.group_name, base_state, acm, ctx_lock);
  if (!gc)
    return false;
  if (b200sv_check_batch(gc->ctx, q, n_states, flags, valid_bitmap) != B200SV_OK)
  {
    RCLCPP_ERROR(getLogger(), "b200sv_check_batch failed: %s", b200sv_last_error(gc->ctx));
    std::fill(valid_bitmap, valid_bitmap + (n_states + 7) / 8, static_cast<uint8_t>(0));
    return false;
  }
  return true;
}

bool CollisionEnvB200::getCollisionGradientsBatch(const CollisionRequest& req, const moveit::core::RobotState& base_state,
                                                  const AllowedCollisionMatrix* acm, const double* q, std::size_t n_states,
                                                  std::vector<std::vector<GradientInfo>>& gradients) const
{
  gradients.clear();
  std::unique_lock<std::mutex> ctx_lock;
  std::shared_ptr<GroupContext> gc = getContext(req.group_name, base_state, acm, ctx_lock);
  if (!gc)
    return false;
  const std::size_t n_entries = gc->entry_names.size();
  const std::size_t n_spheres = gc->sphere_radii.size();
  std::vector<double> dist(n_states * n_spheres), grad(3 * n_states * n_spheres), centres(3 * n_states * n_spheres);
  std::vector<double> closest(n_states * n_entries);
  std::vector<uint8_t> types(n_states * n_spheres);
  if (b200sv_collision_gradients(gc->ctx, q, n_states, dist.data(), grad.data(), types.data(), closest.data(),
                                 centres.data()) != B200SV_OK)
  {
    RCLCPP_ERROR(getLogger(), "b200sv_collision_gradients failed: %s", b200sv_last_error(gc->ctx));
    return false;
  }
  // the reference keeps ONE gradients_ entry per attached body; the context holds one entry per shape of it (b200sv.h:
  // body_id), in order: consecutive entries of one body are appended to the same GradientInfo
  std::vector<std::size_t> out_of(n_entries);
  std::size_t n_out = 0;
  for (std::size_t e = 0; e < n_entries; ++e)
  {
    const bool same_body = e > 0 && gc->entry_is_attached[e] && gc->entry_is_attached[e - 1] && gc->body_id[e] == gc->body_id[e - 1];
    out_of[e] = same_body ? n_out - 1 : n_out++;
  }
  gradients.assign(n_states, std::vector<GradientInfo>(n_out));
  for (std::size_t s = 0; s < n_states; ++s)
    for (std::size_t e = 0; e < n_entries; ++e)
    {
      GradientInfo& g = gradients[s][out_of[e]];
      g.joint_name = gc->entry_joint_names[e];
      g.closest_distance = std::min(g.closest_distance, closest[s * n_entries + e]);
      for (int32_t k = gc->sphere_start[e]; k < gc->sphere_start[e + 1]; ++k)
      {
        const std::size_t at = s * n_spheres + static_cast<std::size_t>(k);
        g.sphere_radii.push_back(gc->sphere_radii[static_cast<std::size_t>(k)]);
        g.sphere_locations.emplace_back(centres[3 * at], centres[3 * at + 1], centres[3 * at + 2]);
        g.distances.push_back(dist[at]);
        g.gradients.emplace_back(grad[3 * at], grad[3 * at + 1], grad[3 * at + 2]);
        g.types.push_back(static_cast<CollisionType>(types[at]));
      }
    }
  return true;
}

void CollisionEnvB200::checkSingle(const CollisionRequest& req, CollisionResult& res,
                                   const moveit::core::RobotState& state, const AllowedCollisionMatrix* acm,
                                   uint32_t flags) const
{
  const moveit::core::JointModelGroup* jmg = robot_model_->getJointModelGroup(req.group_name);
  if (!jmg)
    return;  // the reference logs and checks nothing (:716-720)
  std::vector<double> q(jmg->getVariableCount());
  state.copyJointGroupPositions(jmg, q.data());
  if (req.contacts)
  {  // the contacts the reference would report, in its order and under its limits (:298-339, 571-638, 1589-1629)
    if (res.contact_count >= req.max_contacts)
      return;
    std::unique_lock<std::mutex> ctx_lock;
    std::shared_ptr<GroupContext> gc = getContext(req.group_name, state, acm, ctx_lock);
    if (!gc)
    {
      res.collision = true;  // fail closed: no context (no device, robot too large, CUDA error) is not "collision-free"
      return;
    }
    const uint32_t left = static_cast<uint32_t>(req.max_contacts - res.contact_count);
    std::vector<b200sv_contact> cons(left);
    uint8_t coll = 0;
    int32_t count = 0;
    if (b200sv_check_contacts(gc->ctx, q.data(), 1, flags, left, static_cast<uint32_t>(req.max_contacts_per_pair), &coll,
                              &count, cons.data()) != B200SV_OK)
    {
      RCLCPP_ERROR(getLogger(), "b200sv_check_contacts failed: %s", b200sv_last_error(gc->ctx));
      res.collision = true;  // fail closed
      return;
    }
    res.collision = res.collision || coll;
    for (int32_t k = 0; k < count; ++k)
    {
      Contact con;
      con.pos = Eigen::Vector3d(cons[k].pos[0], cons[k].pos[1], cons[k].pos[2]);
      con.body_name_1 = gc->entry_names[cons[k].body_1];
      con.body_type_1 = gc->entry_is_attached[cons[k].body_1] ? BodyTypes::ROBOT_ATTACHED : BodyTypes::ROBOT_LINK;
      if (cons[k].body_2 >= 0)
      {
        con.body_name_2 = gc->entry_names[cons[k].body_2];
        con.body_type_2 = gc->entry_is_attached[cons[k].body_2] ? BodyTypes::ROBOT_ATTACHED : BodyTypes::ROBOT_LINK;
      }
      else
      {
        con.body_name_2 = cons[k].body_2 == -1 ? "self" : "environment";
        con.body_type_2 = cons[k].body_2 == -1 ? BodyTypes::ROBOT_LINK : BodyTypes::WORLD_OBJECT;
      }
      res.contacts[std::make_pair(con.body_name_1, con.body_name_2)].push_back(con);
      ++res.contact_count;
    }
    return;
  }
  // checkCollisionBatch clears the bit on every failure path, so an engine failure reads as "in collision" (fail closed)
  uint8_t bit = 0;
  checkCollisionBatch(req, state, acm, q.data(), 1, &bit, flags);
  if (!(bit & 1))
    res.collision = true;
}

void CollisionEnvB200::checkSelfCollision(const CollisionRequest& req, CollisionResult& res,
                                          const moveit::core::RobotState& state) const
{
  checkSingle(req, res, state, nullptr, B200SV_CHECK_SELF | B200SV_CHECK_INTRA);
}
void CollisionEnvB200::checkSelfCollision(const CollisionRequest& req, CollisionResult& res,
                                          const moveit::core::RobotState& state, const AllowedCollisionMatrix& acm) const
{
  checkSingle(req, res, state, &acm, B200SV_CHECK_SELF | B200SV_CHECK_INTRA);
}
void CollisionEnvB200::checkRobotCollision(const CollisionRequest& req, CollisionResult& res,
                                           const moveit::core::RobotState& state) const
{
  checkSingle(req, res, state, nullptr, B200SV_CHECK_WORLD);
}
void CollisionEnvB200::checkRobotCollision(const CollisionRequest& req, CollisionResult& res,
                                           const moveit::core::RobotState& state, const AllowedCollisionMatrix& acm) const
{
  checkSingle(req, res, state, &acm, B200SV_CHECK_WORLD);  // the world check ignores the ACM (:1561-1643)
}
void CollisionEnvB200::checkRobotCollision(const CollisionRequest&, CollisionResult&, const moveit::core::RobotState&,
                                           const moveit::core::RobotState&) const
{
  RCLCPP_ERROR(getLogger(), "Continuous collision checking not implemented");  // as the reference (:1502-1515)
}
void CollisionEnvB200::checkRobotCollision(const CollisionRequest&, CollisionResult&, const moveit::core::RobotState&,
                                           const moveit::core::RobotState&, const AllowedCollisionMatrix&) const
{
  RCLCPP_ERROR(getLogger(), "Continuous collision checking not implemented");
}
void CollisionEnvB200::checkCollision(const CollisionRequest& req, CollisionResult& res,
                                      const moveit::core::RobotState& state) const
{
  checkSingle(req, res, state, nullptr, B200SV_CHECK_ALL);
}
void CollisionEnvB200::checkCollision(const CollisionRequest& req, CollisionResult& res,
                                      const moveit::core::RobotState& state, const AllowedCollisionMatrix& acm) const
{
  checkSingle(req, res, state, &acm, B200SV_CHECK_ALL);
}
}  // namespace collision_detection
