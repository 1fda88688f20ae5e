// renderer.hpp -- the Renderer seam of the reference, signature for signature
// (gpugi/renderer/renderer.hpp:25-51).  This abstract class IS the drop-in boundary.
#pragma once

#include <memory>
#include <string>

namespace gl
{
	class Texture2D;
	class TextureCubemap;
}

class Scene;
class Camera;
class RendererSystem;

/// Abstract base class for all renderers; unified interface for interop with RendererSystem.
class Renderer
{
public:
	virtual ~Renderer() = default;

	/// Returns a unique name for this renderer type ("PT", "LPT", "BPT").
	virtual std::string GetName() const = 0;

	virtual void SetScene(std::shared_ptr<Scene> _scene) {}
	virtual void SetCamera(const Camera& camera) {}
	virtual void SetScreenSize(const gl::Texture2D& _newBackbuffer) = 0;
	virtual void SetEnvironmentMap(std::shared_ptr<gl::TextureCubemap> _envMap) {}
	/// Draw command: one sample pass into the backbuffer.
	virtual void Draw() = 0;

	const RendererSystem& GetRendererSystem() const { return m_rendererSystem; }

protected:
	Renderer(RendererSystem& _rendererSystem) : m_rendererSystem(_rendererSystem) {}

	RendererSystem& m_rendererSystem;
};
